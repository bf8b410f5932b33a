"""Parameter ensembles: M members of one model grid as ONE device model.

The reference runs an ensemble (calibration: wflow/wflow_fit.py:153-158 multiplies M, KsatVer, N, ... of an
area by trial factors; :235-257, :315-347 rerun the whole model per trial) as M serial runs of the same
network.  Members never interact, so here their grids sit side by side in one (nrow, M * ncol) mosaic: every
member is a set of independent basins of the drainage forest, the storage layout packs them into groups like
any other basins, and one step of the model advances all members -- each level of the
sweep carries M times the work.  Members are sharded across GPUs with parallel.shard_members.

Results per member are bit-identical to a stand-alone SbmModel with that member's parameters
(tests/test_gpu_ensemble.py): the same kernels run on the same cells in the same order.
"""
import numpy as np

from . import core

_DCOL = {1: -1, 4: -1, 7: -1, 3: 1, 6: 1, 9: 1}


def _member_ldd(ldd):
    """The member grid's LDD as it may be tiled: a cell whose flow leaves the grid through the west / east edge
    would enter the neighbouring member, so it is taken out -- set_dd never reaches it from a pit anyway, nor
    anything upstream of it (wflow/wflow_funcs.py:79-95).  The cell set_dd drops by its flat-index-0 quirk
    (:88) is flat index 0 of EVERY member's own run, so it is taken out of every tile as well."""
    ldd = np.ascontiguousarray(ldd, dtype=np.uint8).copy()
    ncol = ldd.shape[1]
    for code, dc in _DCOL.items():
        col = 0 if dc < 0 else ncol - 1
        ldd[ldd[:, col] == code, col] = 0
    if ldd[0, 0] not in (0, 5):
        alone = core.Schedule(ldd)
        order, _, _, _ = alone.partition(1)
        if 0 not in set(order[: alone.num_cells].tolist()):
            ldd[0, 0] = 0
    return ldd


class SbmEnsemble:
    """`members` copies of the grid (ldd, river) in one SbmModel.  Rasters are set for all members at once
    (set) or per member (set_members: array [members, nrow, ncol]); get returns [members, nrow, ncol]."""

    def __init__(self, ldd, river, members, **model_kw):
        self.members = int(members)
        if self.members < 1:
            raise ValueError("an ensemble needs at least one member")
        self.grid_shape = tuple(np.asarray(ldd).shape)
        tile = _member_ldd(ldd)
        riv = np.ascontiguousarray(np.asarray(river) != 0, dtype=np.uint8)
        self.ldd = tile
        self.model = core.SbmModel(np.tile(tile, (1, self.members)), np.tile(riv, (1, self.members)), **model_kw)

    # -- rasters ---------------------------------------------------------------------------
    def _mosaic(self, per_member):
        a = np.asarray(per_member, dtype=np.float32)
        if a.shape != (self.members,) + self.grid_shape:
            raise ValueError("expected an array of shape %s" % ((self.members,) + self.grid_shape,))
        return np.ascontiguousarray(np.concatenate(list(a), axis=1))

    def set(self, name, value):
        """The same raster (or scalar) for every member."""
        a = np.asarray(value, dtype=np.float32)
        if a.ndim == 0 or a.size == 1:
            self.model.set(name, a)
        else:
            self.model.set(name, np.tile(a, (1, self.members)))

    def set_members(self, name, per_member):
        self.model.set(name, self._mosaic(per_member))

    def scale(self, name, base, factors):
        """Member m gets base * factors[m] (float32, like a wflow_fit trial multiplying a table parameter)."""
        base = np.broadcast_to(np.asarray(base, dtype=np.float32), self.grid_shape)
        f = np.asarray(factors, dtype=np.float32).reshape(self.members, 1, 1)
        self.set_members(name, base[None] * f)

    def get(self, name, mv=-999.0):
        a = self.model.get(name, mv)
        nc = self.grid_shape[1]
        return np.stack([a[:, m * nc:(m + 1) * nc] for m in range(self.members)])

    # -- stepping --------------------------------------------------------------------------
    def step(self, nsteps=1):
        self.model.step(nsteps)

    def gauges_create(self, flat_index):
        """Gauge set of the given cells (flat indices of the member grid) in every member; sampling returns
        [members * len(flat_index)] values, member-major."""
        idx = np.asarray(flat_index, dtype=np.int64)
        nrow, nc = self.grid_shape
        r, c = idx // nc, idx % nc
        allidx = np.concatenate([r * (nc * self.members) + m * nc + c for m in range(self.members)])
        return self.model.gauges_create(allidx), idx.size

    def gauges_sample(self, handle, field, mv=-999.0):
        h, n = handle
        out = np.zeros(self.members * n, np.float32)
        self.model.gauges_sample(h, field, out.ctypes.data, mv=mv, wait=True)
        return out.reshape(self.members, n)

    def close(self):
        self.model.close()
