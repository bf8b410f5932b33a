"""N > 1 path on CPU: basin partitioning and the gauge gather over torch.distributed (gloo, world 2)."""
import os
import subprocess
import sys

import numpy as np

from wflow_b200 import parallel, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_basin_labels_and_partition():
    ldd = synth.ldd_basin_tiles(96, 128, tile=32)  # 3 x 4 = 12 independent basins
    labels, sizes = parallel.basin_labels(ldd)
    assert sizes.size == 12 and sizes.sum() == (labels > 0).sum()
    # every cell carries the label of the pit it reaches by following the ldd
    pits = np.argwhere(ldd == 5)
    for i, (r, c) in enumerate(pits):
        assert labels[r, c] == i + 1
        assert (labels[(r // 32) * 32:(r // 32 + 1) * 32, (c // 32) * 32:(c // 32 + 1) * 32] == i + 1).sum() == sizes[i]
    parts = parallel.partition(sizes, 5)
    allb = np.sort(np.concatenate(parts))
    assert np.array_equal(allb, np.arange(12))          # every basin exactly once
    loads = [sizes[p].sum() for p in parts]
    assert max(loads) - min(loads) <= sizes.max()        # LPT balance bound
    shard = parallel.shard_ldd(ldd, labels, parts[0])
    assert (shard > 0).sum() == sizes[parts[0]].sum()
    assert [len(parallel.shard_members(256, r, 8)) for r in range(8)] == [32] * 8
    assert sum(len(parallel.shard_members(10, r, 4)) for r in range(4)) == 10


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
from wflow_b200 import parallel, synth
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ldd = synth.ldd_basin_tiles(64, 96, tile=32)             # 6 basins, the same grid on every rank
labels, sizes = parallel.basin_labels(ldd)
mine = parallel.partition(sizes, world)[rank]
shard = parallel.shard_ldd(ldd, labels, mine)
# stand-in for the per-rank outlet discharges: one value per own basin (its cell count)
series = np.array([float((labels == b + 1).sum()) for b in mine], np.float32)
parts = parallel.gather_series(series)
if rank == 0:
    assert len(parts) == world
    got = np.sort(np.concatenate(parts))
    assert np.array_equal(got, np.sort(sizes.astype(np.float32))), (got, sizes)
    own = parallel.partition(sizes, world)
    for r in range(world):
        assert np.array_equal(parts[r], sizes[own[r]].astype(np.float32))
    print("GATHER_OK", world, int((shard > 0).sum()))
dist.destroy_process_group()
'''


def test_gather_over_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GATHER_OK 2" in res.stdout


HBV_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %r)
from wflow_b200 import parallel, synth
from oracle import hbv_oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 48
ldd, static, state = synth.hbv_case(n, n, kind="tiles", tile=16, seed=5)      # 9 basins, the same case on every rank
state["SurfaceRunoff"][0, 0] = 0.0
labels, sizes = parallel.basin_labels(ldd)
mine = parallel.partition(sizes, world)[rank]
shard = parallel.shard_ldd(ldd, labels, mine)
fields = dict(static); fields.update(state)
forc = synth.Forcing(n, n, seed=5, rain_mean=14.0, t_mean=3.0)
m = hbv_oracle.HbvOracle(shard, fields)                                         # this rank's basins only
whole = hbv_oracle.HbvOracle(ldd, fields) if rank == 0 else None
for t in range(3):
    P, PET, T = forc.step(t)
    m.step(P, PET, T)
    if whole is not None:
        whole.step(P, PET, T)
pits = np.argwhere(ldd == 5)
own = [tuple(pits[b]) for b in mine]
series = np.array([m.get("SurfaceRunoff")[r, c] for r, c in own], np.float32)   # outlet discharge of the own basins
parts = parallel.gather_series(series)
if rank == 0:
    parts_of = parallel.partition(sizes, world)
    for r in range(world):
        want = np.array([whole.get("SurfaceRunoff")[tuple(pits[b])] for b in parts_of[r]], np.float32)
        assert np.array_equal(parts[r], want), (r, parts[r], want)       # a basin computed alone = the basin inside the whole grid
    assert float(np.concatenate(parts).max()) > 0.0
    print("HBV_SHARDS_OK", world, sum(len(p) for p in parts))
dist.destroy_process_group()
'''


def test_hbv_shards_by_basin_over_gloo_world_size_2(tmp_path):
    """wflow_hbv shards like wflow_sbm (independent basins per rank, gather of the outlets): every rank steps its basins with
    the CPU restatement, the gathered outlet discharges are bit-identical to the whole grid's."""
    script = tmp_path / "hbv_worker.py"
    script.write_text(HBV_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29519", str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "HBV_SHARDS_OK 2 9" in res.stdout
