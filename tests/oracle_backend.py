"""TEST INFRASTRUCTURE: the CPU oracle behind the interface of wflow_b200.core.SbmModel, so that the
host side of the framework mirror (ini -> parameter maps -> time loop -> csv/tss outputs) can be
driven on a machine without a GPU and the ORACLE itself can be pinned against the numbers the
reference's own acceptance tests hold (tests/test_wflow_sbm.py:60-68,
tests/test_wflow_sbm_example.py:54).  Tests swap it in with monkeypatch; the product never sees it
(tests/test_cabi.py::test_product_does_not_import_the_oracle)."""
import numpy as np

from oracle import oracle as O

F32 = np.float32


class OracleSbmModel:
    def __init__(self, ldd, river=None, *, schedule=None, timestepsecs=86400, layer_thickness=(), model_snow=1,
                 soil_inf_redu=0, transfer_method=0, ust=0, glacier=0, it_kin_l=1, it_kin_r=1, device=0, beta=0.0):
        self.shape = ldd.shape
        self._o = O.Oracle(ldd, river, timestepsecs=timestepsecs, layer_thickness=layer_thickness,
                           modelSnow=model_snow, soilInfRedu=soil_inf_redu, transferMethod=transfer_method, ust=ust,
                           it_kinL=it_kin_l, it_kinR=it_kin_r, glacier=glacier)
        self.max_layers = self._o.maxLayers
        self._o.set("River", (np.asarray(river) != 0).astype(F32) if river is not None else np.zeros(self.shape, F32))
        self._o.set("Beta", F32(0.6))
        self._mask = np.zeros(self.shape, bool)
        self._mask.flat[self._o.net.flat_nodes] = True
        self._areas = []
        O.set_threads(8)

    num_cells = property(lambda self: int(self._o.net.flat_nodes.size))
    num_levels = property(lambda self: len(self._o.net.lev_off) - 1)
    num_river_cells = property(lambda self: int(self._o.rnet.flat_nodes.size))

    def close(self):
        pass

    def set(self, name, value):
        self._o.set(name, value)

    def get(self, name, mv=-999.0):
        return np.where(self._mask, self._o[name], F32(mv)).astype(F32)

    def request(self, *names, enable=True):
        for n in names:
            if n not in self._o.names:
                raise KeyError(n)
        self._o.want(*names)

    def device_ptr(self, name):
        if name not in self._o.names:
            raise KeyError(name)
        return 0, 0

    def estimate_iterations(self, which):
        return self._o.estimate_iterations(which)

    def set_substeps(self, it_l, it_r):
        self._o.set_substeps(it_l, it_r)

    def step(self, nsteps=1):
        for _ in range(nsteps):
            self._o.step()

    def synchronize(self):
        pass

    def set_option(self, name, value):
        """[model] options of the optional modules: the oracle restates none of them (a value that switches one on is refused)."""
        if name in ("MassWasting", "Abstractions") and value:
            raise NotImplementedError("the CPU oracle does not restate %s" % name)

    def irrigation_create(self, areas, intakes):
        raise NotImplementedError("the CPU oracle does not restate the irrigation module")

    def quick_suspend(self):
        self._saved = {k: v.copy() for k, v in self._o.fields.items()}

    def quick_resume(self):
        for k, v in self._saved.items():
            self._o.fields[k][...] = v

    # wf_OutputTimeSeriesArea.writestep (wf_DynamicFramework.py:446-499): area statistic over the id map
    def area_create(self, idmap):
        ids = np.where(self._mask, np.asarray(idmap, np.int32), 0)
        uniq = np.unique(ids[ids > 0]).astype(np.int32)
        self._areas.append((ids, uniq))
        return len(self._areas) - 1, uniq

    def area_reduce(self, area, field, op="average"):
        ids, uniq = self._areas[area]
        a = self._o[field].astype(np.float64)
        fn = {"average": np.mean, "total": np.sum, "maximum": np.max, "minimum": np.min}[op]
        return np.array([fn(a[ids == i]) for i in uniq], F32)

    def sample(self, field, flat_index, mv=-999.0):
        return self.get(field, mv).reshape(-1)[np.asarray(flat_index, np.int64)]


class OracleHbvModel:
    """The float32-numpy restatement of wflow_hbv's timestep (oracle/hbv_oracle.py) behind the interface of
    wflow_b200.hbv.HbvModel: the host side of framework_hbv runs without a GPU, and the restatement is pinned against the
    reference's own acceptance numbers (tests/test_wflow_hbv.py of the reference)."""

    def __init__(self, ldd, *, timestepsecs=86400, kinwaveIters=0, kinwaveTstep=0, set_kquick_flow=0, external_qbase=0,
                 mass_wasting=0, glacier=0, device=0, beta=0.0):
        self.shape = ldd.shape
        self._ldd = np.asarray(ldd)
        self._kw = dict(timestepsecs=timestepsecs, kinwaveIters=kinwaveIters, kinwaveTstep=kinwaveTstep,
                        setKquickFlow=set_kquick_flow, externalQbase=external_qbase, massWasting=mass_wasting)
        self._fields, self._o, self._areas = {}, None, []
        net = O.Network(self._ldd)
        self._mask = np.zeros(self.shape, bool)
        self._mask.flat[net.flat_nodes] = True
        self.it_kin = 1
        O.set_threads(8)

    def close(self):
        pass

    def set(self, name, value):
        v = np.array(np.broadcast_to(np.asarray(value, F32), self.shape))
        if self._o is not None:
            self._o.set(name, v)
        else:
            self._fields[name] = v

    def set_many(self, fields):
        for k, v in fields.items():
            self.set(k, v)

    def _model(self):
        if self._o is None:
            from oracle import hbv_oracle
            self._o = hbv_oracle.HbvOracle(self._ldd, self._fields, **self._kw)
        return self._o

    def get(self, name, mv=-999.0):
        return np.where(self._mask, self._model().get(name), F32(mv)).astype(F32)

    def request(self, *names, enable=True):
        pass

    def device_ptr(self, name):
        return 0, 0

    def dynamic(self, P, PET, TEMP=None, Inflow=None, Seepage=None):
        o = self._model()
        o.step(P, PET, TEMP, Inflow=Inflow, Seepage=Seepage)
        self.it_kin = o.last_it

    def quick_suspend(self):
        pass

    def quick_resume(self):
        raise NotImplementedError

    def area_create(self, idmap):
        ids = np.where(self._mask, np.asarray(idmap, np.int32), 0)
        uniq = np.unique(ids[ids > 0]).astype(np.int32)
        self._areas.append((ids, uniq))
        return len(self._areas) - 1, uniq

    def area_reduce(self, area, field, op="average"):
        ids, uniq = self._areas[area]
        a = self._model().get(field).astype(np.float64)
        fn = {"average": np.mean, "total": np.sum, "maximum": np.max, "minimum": np.min}[op]
        return np.array([fn(a[ids == i]) for i in uniq], F32)
