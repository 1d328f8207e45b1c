from .model import Model, load_model  # noqa: F401
from .mps import MatrixProductState, MPS_initialize, TensorNetwork  # noqa: F401
from .smpo import SMPO_initialize, SpacedMatrixProductOperator  # noqa: F401
from .mpo import MatrixProductOperator, MPO_initialize  # noqa: F401
