"""tn4ml_b200: the data-parallel hot path of tn4ml (batched contraction of embedded product
states through MPS / MPO / SMPO chains, forward and backward) on hand-written sm_100a kernels.

Public names mirror ``tn4ml`` (reference: tn4ml/__init__.py)."""
from . import embeddings, initializers, metrics, models, optim, strategy, util  # noqa: F401
from .embeddings import (  # noqa: F401
    Embedding, FourierEmbedding, GaussianRBFEmbedding, HermiteEmbedding, JaxArraysEmbedding, LaguerreEmbedding,
    LegendreEmbedding, LinearComplementEmbedding, PolynomialEmbedding, TrigonometricEmbedding)
from .models import (  # noqa: F401
    Model, MatrixProductOperator, MatrixProductState, MPO_initialize, MPS_initialize, SMPO_initialize,
    SpacedMatrixProductOperator, TensorNetwork, load_model)
from .strategy import Strategy, Sweeps  # noqa: F401
from .util import EarlyStopping, TrainingType  # noqa: F401

__version__ = "0.1.0"
