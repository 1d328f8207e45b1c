"""A small named-index tensor network with the slice of quimb.tensor's API that tn4ml's hot path touches.

Semantics follow quimb's documentation: a Tensor is (data, inds, tags); combining networks with `&` / `|` keeps index names, so
tensors sharing a name are contracted over it by `^ all`; `.H` conjugates; `norm()` is sqrt(<tn|tn>) with the bra's INNER
indices renamed; `MatrixProductState(arrays, shape="lrp", site_ind_id="k{}", site_tag_id="I{}")` gives site i the physical index
`k{i}`, open-boundary end tensors of rank 2 (first: (r, p), last: (l, p)), and bond names unique to the instance.
"""
import itertools
import uuid

import numpy as np

from . import tensor_core, tensor_1d, tensor_arbgeom  # noqa: F401,E402  (filled below)


def rand_uuid():
    return "_" + uuid.uuid4().hex[:10]


class Tensor:
    def __init__(self, data=1.0, inds=(), tags=None, left_inds=None):
        if isinstance(data, Tensor):
            inds, tags, data = data.inds, data.tags, data.data
        self._data = np.asarray(data)
        self._inds = tuple(inds)
        if isinstance(tags, str):
            tags = [tags]
        self._tags = list(dict.fromkeys(tags or []))
        if self._data.ndim != len(self._inds):
            raise ValueError(f"data of rank {self._data.ndim} with {len(self._inds)} indices {self._inds}")

    data = property(lambda s: s._data)
    inds = property(lambda s: s._inds)
    tags = property(lambda s: s._tags)
    shape = property(lambda s: s._data.shape)
    ndim = property(lambda s: s._data.ndim)
    dtype = property(lambda s: s._data.dtype)

    def modify(self, data=None, inds=None, tags=None, **_kw):
        if data is not None:
            self._data = np.asarray(data)
        if inds is not None:
            self._inds = tuple(inds)
        if tags is not None:
            self._tags = list(dict.fromkeys([tags] if isinstance(tags, str) else tags))

    def copy(self, deep=False):
        return Tensor(self._data.copy() if deep else self._data, self._inds, list(self._tags))

    def conj(self):
        return Tensor(np.conj(self._data), self._inds, list(self._tags))

    H = property(conj)

    def reindex(self, index_map, inplace=False):
        t = self if inplace else self.copy()
        t._inds = tuple(index_map.get(i, i) for i in t._inds)
        return t

    reindex_ = lambda self, m: self.reindex(m, inplace=True)  # noqa: E731

    def transpose(self, *inds, inplace=False):
        perm = [self._inds.index(i) for i in inds]
        t = self if inplace else self.copy()
        t._data, t._inds = np.transpose(self._data, perm), tuple(inds)
        return t

    def bonds(self, other):
        return tuple(i for i in self._inds if i in other.inds)

    def ind_size(self, ind):
        return self._data.shape[self._inds.index(ind)]

    def norm(self):
        return np.linalg.norm(self._data)

    def squeeze(self, inplace=False):
        keep = [k for k, d in enumerate(self.shape) if d != 1]
        t = self if inplace else self.copy()
        t._data, t._inds = self._data.reshape([self.shape[k] for k in keep]), tuple(self._inds[k] for k in keep)
        return t

    def add_tag(self, tag):
        if tag not in self._tags:
            self._tags.append(tag)

    def drop_tags(self, tags=None):
        tags = [tags] if isinstance(tags, str) else list(tags or self._tags)
        self._tags = [t for t in self._tags if t not in tags]

    # numpy interop: jnp.linalg.norm(tensor), tensor / scalar, jnp.squeeze(tensor.data) ...
    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._data, dtype=dtype)

    def _bin(self, other, op):
        o = other.data if isinstance(other, Tensor) else other
        return Tensor(op(self._data, o), self._inds, list(self._tags))

    __truediv__ = lambda s, o: s._bin(o, np.divide)  # noqa: E731
    __mul__ = lambda s, o: s._bin(o, np.multiply)  # noqa: E731
    __rmul__ = __mul__
    __add__ = lambda s, o: s._bin(o, np.add)  # noqa: E731
    __sub__ = lambda s, o: s._bin(o, np.subtract)  # noqa: E731
    __neg__ = lambda s: Tensor(-s._data, s._inds, list(s._tags))  # noqa: E731
    __pow__ = lambda s, o: s._bin(o, np.power)  # noqa: E731

    def __and__(self, other):
        return TensorNetwork([self, other])

    __or__ = __and__

    def __repr__(self):
        return f"Tensor(shape={self.shape}, inds={self._inds}, tags={self._tags})"


def _contract_pair(a, b):
    shared = [i for i in a.inds if i in b.inds]
    ax_a = [a.inds.index(i) for i in shared]
    ax_b = [b.inds.index(i) for i in shared]
    data = np.tensordot(a.data, b.data, axes=(ax_a, ax_b))
    inds = tuple(i for i in a.inds if i not in shared) + tuple(i for i in b.inds if i not in shared)
    return Tensor(data, inds, list(dict.fromkeys(list(a.tags) + list(b.tags))))


def tensor_contract(*tensors, output_inds=None, **_kw):
    """Sum over every index name shared by two tensors; greedy pairwise order (smallest result first)."""
    ts = [t for t in tensors]
    if not ts:
        return 1.0
    # indices appearing more than twice would be hyper-indices: not used by tn4ml
    while len(ts) > 1:
        best = None
        for i, j in itertools.combinations(range(len(ts)), 2):
            shared = [x for x in ts[i].inds if x in ts[j].inds]
            if not shared:
                continue
            size = 1
            for x, d in zip(ts[i].inds + ts[j].inds, ts[i].shape + ts[j].shape):
                if x not in shared:
                    size *= d
            if best is None or size < best[0]:
                best = (size, i, j)
        if best is None:  # disconnected: outer product of the first two
            i, j = 0, 1
        else:
            _, i, j = best
        c = _contract_pair(ts[i], ts[j])
        ts = [t for k, t in enumerate(ts) if k not in (i, j)] + [c]
    t = ts[0]
    if output_inds is not None and tuple(output_inds) != t.inds:
        t = t.transpose(*output_inds)
    if t.ndim == 0:
        return t.data[()]
    return t


class TensorNetwork:
    _EXTRA_PROPS = ()

    def __init__(self, ts=(), *, virtual=False, check_collisions=True, **_kw):
        self._tensors = []
        if isinstance(ts, TensorNetwork):
            ts = ts.tensors
        for t in ts:
            if isinstance(t, TensorNetwork):
                self._tensors.extend(t.tensors if virtual else [x.copy() for x in t.tensors])
            else:
                self._tensors.append(t if virtual else t.copy())

    tensors = property(lambda s: tuple(s._tensors))
    arrays = property(lambda s: tuple(t.data for t in s._tensors))
    num_tensors = property(lambda s: len(s._tensors))

    def __iter__(self):
        return iter(self._tensors)

    def __len__(self):
        return len(self._tensors)

    def copy(self, virtual=False, deep=False):
        new = object.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        new._tensors = [t if virtual else t.copy(deep=deep) for t in self._tensors]
        return new

    def conj(self):
        new = self.copy()
        new._tensors = [t.conj() for t in self._tensors]
        return new

    H = property(conj)

    def _combine(self, other, virtual):
        if isinstance(other, TensorNetwork):
            # quimb (check_collisions): inner indices of the second network that are also inner indices of the first are
            # renamed -- so that `tn.H & tn` is the overlap and not a sum over four-fold repeated bonds
            clash = set(self.inner_inds()) & set(other.inner_inds())
            if clash:
                other = other.copy()
                other.reindex({i: rand_uuid() for i in clash}, inplace=True)
            others = other.tensors
        else:
            others = [other]
        return TensorNetwork(list(self._tensors) + list(others), virtual=virtual)

    def __and__(self, other):
        return self._combine(other, virtual=False)

    def __or__(self, other):
        return self._combine(other, virtual=True)

    def ind_map(self):
        m = {}
        for k, t in enumerate(self._tensors):
            for i in t.inds:
                m.setdefault(i, []).append(k)
        return m

    def outer_inds(self):
        return tuple(i for i, v in self.ind_map().items() if len(v) == 1)

    def inner_inds(self):
        return tuple(i for i, v in self.ind_map().items() if len(v) > 1)

    def contract(self, tags=..., output_inds=None, **_kw):
        return tensor_contract(*self._tensors, output_inds=output_inds)

    def __xor__(self, what):
        if what is all or what is Ellipsis:
            return self.contract()
        raise NotImplementedError("only `^ all` is supported by the stand-in")

    def contract_ind(self, ind, **_kw):
        if isinstance(ind, (list, tuple)):  # several indices: the tensors carrying any of them are contracted together
            for i in ind:
                if i in self.ind_map():
                    self.contract_ind(i)
            return
        ks = self.ind_map().get(ind, [])
        if len(ks) == 1:  # an outer index: quimb contracts the tensor with output_inds = its other indices, i.e. sums over it
            t = self._tensors[ks[0]]
            ax = t.inds.index(ind)
            t.modify(data=t.data.sum(axis=ax), inds=tuple(i for i in t.inds if i != ind))
            return
        if len(ks) != 2:
            raise ValueError(f"index {ind} is on {len(ks)} tensors")
        c = _contract_pair(self._tensors[ks[0]], self._tensors[ks[1]])
        self._tensors = [t for k, t in enumerate(self._tensors) if k not in ks]
        self._tensors.insert(ks[0], c)

    def _find_by_tags(self, tags):
        tags = [tags] if isinstance(tags, str) else list(tags)
        hits = [k for k, t in enumerate(self._tensors) if all(tg in t.tags for tg in tags)]
        if len(hits) != 1:
            raise KeyError(f"{len(hits)} tensors carry the tags {tags}")
        return hits[0]

    def contract_between(self, tags1, tags2, **_kw):
        """Contract the two tensors identified by tags1 / tags2 (over every index they share); the result replaces them."""
        i, j = self._find_by_tags(tags1), self._find_by_tags(tags2)
        c = _contract_pair(self._tensors[i], self._tensors[j])
        self._tensors = [t for k, t in enumerate(self._tensors) if k not in (i, j)]
        self._tensors.insert(min(i, j), c)

    def drop_tags(self, tags=None):
        for t in self._tensors:
            t.drop_tags(tags)

    def fuse_multibonds(self, inplace=False, **_kw):
        """Where two tensors share several indices, fuse those into one index."""
        tn = self if inplace else self.copy()
        ts = tn._tensors
        for i in range(len(ts)):
            for j in range(i + 1, len(ts)):
                shared = [x for x in ts[i].inds if x in ts[j].inds]
                if len(shared) < 2:
                    continue
                new = rand_uuid()
                for t in (ts[i], ts[j]):
                    rest = [x for x in t.inds if x not in shared]
                    tt = t.transpose(*(rest + shared))
                    n = int(np.prod([t.ind_size(x) for x in shared]))
                    t.modify(data=tt.data.reshape(tt.shape[:len(rest)] + (n,)), inds=tuple(rest) + (new,))
        return tn

    fuse_multibonds_ = lambda self, **kw: self.fuse_multibonds(inplace=True, **kw)  # noqa: E731

    def reindex(self, index_map, inplace=False):
        tn = self if inplace else self.copy()
        for t in tn._tensors:
            t.reindex(index_map, inplace=True)
        return tn

    reindex_ = lambda self, m: self.reindex(m, inplace=True)  # noqa: E731

    def mangle_inner_(self):
        m = {i: rand_uuid() for i in self.inner_inds()}
        return self.reindex(m, inplace=True)

    def norm(self, **_kw):
        bra = self.conj()
        bra.mangle_inner_()
        n2 = tensor_contract(*(list(self._tensors) + list(bra._tensors)))
        return np.sqrt(np.abs(n2))

    def multiply(self, x, inplace=False, spread_over=8):
        tn = self if inplace else self.copy()
        tn._tensors[0].modify(data=tn._tensors[0].data * x)
        return tn

    def __truediv__(self, x):
        return self.multiply(1.0 / x)

    def normalize(self, inplace=True):
        n = self.norm()
        return self.multiply(1.0 / n, inplace=inplace)


tensor_core.Tensor = Tensor
tensor_core.TensorNetwork = TensorNetwork
tensor_core.tensor_contract = tensor_contract
tensor_core.rand_uuid = rand_uuid
tensor_core.get_tags = lambda tn: list(dict.fromkeys(t_ for t in tn for t_ in t.tags))


class TensorNetworkGen(TensorNetwork):
    pass


class TensorNetwork1D(TensorNetwork):
    _EXTRA_PROPS = ("_site_tag_id", "_L")

    L = property(lambda s: s._L)
    nsites = property(lambda s: s._L)
    site_tag_id = property(lambda s: s._site_tag_id)
    sites = property(lambda s: tuple(range(s._L)))

    def site_tag(self, i):
        return self._site_tag_id.format(i % self._L)

    site_tags = property(lambda s: tuple(s.site_tag(i) for i in range(s._L)))

    def __getitem__(self, i):
        if isinstance(i, int):
            tag = self.site_tag(i)
            for t in self._tensors:
                if tag in t.tags:
                    return t
            raise KeyError(tag)
        for t in self._tensors:
            if i in t.tags:
                return t
        raise KeyError(i)


class TensorNetwork1DVector(TensorNetwork1D):
    site_ind_id = property(lambda s: s._site_ind_id)

    def site_ind(self, i):
        return self._site_ind_id.format(i % self._L)


class TensorNetwork1DOperator(TensorNetwork1D):
    """Upper / lower site indices named by format strings; assigning a new format string renames the indices (quimb's
    documented behaviour of `upper_ind_id` / `lower_ind_id`)."""
    _NDIMS = 1

    def upper_ind(self, i):
        return self._upper_ind_id.format(i)

    def lower_ind(self, i):
        return self._lower_ind_id.format(i)

    def _rename(self, old_id, new_id):
        if old_id == new_id:
            return
        self.reindex({old_id.format(i): new_id.format(i) for i in range(self._L)}, inplace=True)

    def _get_upper(self):
        return self._upper_ind_id

    def _set_upper(self, new):
        self._rename(self._upper_ind_id, new)
        self._upper_ind_id = new

    def _get_lower(self):
        return self._lower_ind_id

    def _set_lower(self, new):
        self._rename(self._lower_ind_id, new)
        self._lower_ind_id = new

    upper_ind_id = property(_get_upper, _set_upper)
    lower_ind_id = property(_get_lower, _set_lower)


class TensorNetwork1DFlat(TensorNetwork1D):
    def bond(self, i, j):
        (b,) = self[i].bonds(self[j])
        return b

    def left_canonize_site(self, i, bra=None):
        """QR of site i (everything but the bond to site i+1 as rows); R is absorbed into site i+1."""
        a, b = self[i], self[i + 1]
        bond = self.bond(i, i + 1)
        k = a.inds.index(bond)
        mat = np.moveaxis(a.data, k, -1)
        shp = mat.shape
        q, r = np.linalg.qr(mat.reshape(-1, shp[-1]))
        a.modify(data=np.moveaxis(q.reshape(shp[:-1] + (q.shape[1],)), -1, k))
        kb = b.inds.index(bond)
        b.modify(data=np.moveaxis(np.tensordot(r, b.data, axes=(1, kb)), 0, kb))

    def canonicalize(self, where, **_kw):
        raise NotImplementedError


class MatrixProductState(TensorNetwork1DVector, TensorNetwork1DFlat):
    _EXTRA_PROPS = ("_site_tag_id", "_site_ind_id", "cyclic", "_L")

    def __init__(self, arrays, *, shape="lrp", tags=None, site_ind_id="k{}", site_tag_id="I{}", **tn_opts):
        if isinstance(arrays, MatrixProductState):
            self.__dict__.update(arrays.copy().__dict__)
            return
        arrays = tuple(arrays)
        self._L = len(arrays)
        self._site_ind_id, self._site_tag_id = site_ind_id, site_tag_id
        self.cyclic = np.ndim(arrays[0]) == 3
        L = self._L
        bonds = [rand_uuid() for _ in range(L)]  # bonds[i]: between site i and i+1 (bonds[L-1] closes a cyclic chain)
        ts = []
        for i, a in enumerate(arrays):
            a = np.asarray(a)
            names = {"l": bonds[(i - 1) % L], "r": bonds[i], "p": site_ind_id.format(i)}
            order = shape
            if not self.cyclic and i == 0:
                order = order.replace("l", "")
            if not self.cyclic and i == L - 1:
                order = order.replace("r", "")
            if L == 1 and not self.cyclic:
                order = "p"
            ts.append(Tensor(a, tuple(names[c] for c in order), [site_tag_id.format(i)] + list([tags] if isinstance(tags, str) else (tags or []))))
        TensorNetwork.__init__(self, ts, virtual=True)


class MatrixProductOperator(TensorNetwork1DOperator, TensorNetwork1DFlat):
    def __init__(self, *a, **k):
        raise NotImplementedError("MatrixProductOperator is not part of the stand-in")


for _name in ("TensorNetwork1D", "TensorNetwork1DVector", "TensorNetwork1DOperator", "TensorNetwork1DFlat", "MatrixProductState",
              "MatrixProductOperator", "TensorNetwork"):
    setattr(tensor_1d, _name, globals()[_name])
tensor_arbgeom.get_coordinate_formatter = lambda ndims: "{}"


class _ArrayOps:
    ndim = staticmethod(np.ndim)


array_ops = _ArrayOps()
