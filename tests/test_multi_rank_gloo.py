"""N > 1 host logic on CPU: world_size-2 gloo run of the point-range-sharded commit.  The GPU context is
replaced by an oracle-backed stand-in (tests may use the oracle as the checker); what is under test is
keaki_b200.dist: shard ranges, partial packing, all_gather order, final combine."""
import os
import random
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from keaki_b200 import dist as kd
from oracle import bn254 as bn
from tests import limbs as L


class OracleCtx:
    """msm_g1 / g1_sum with the Context signature, computed by the big-int oracle."""

    def __init__(self, srs_points):
        self.srs = srs_points

    def msm_g1(self, scalars, n=None, first=0):
        n = scalars.shape[0] if n is None else n
        s = L.fr_vec_from(np.ascontiguousarray(scalars).reshape(-1))[:n]
        p = bn.g1_msm(self.srs[first:first + n], s)
        return L.g1_m(p), 1 if p is None else 0

    def g1_sum(self, pts_xy, inf=None):
        acc = None
        for i in range(pts_xy.shape[0]):
            acc = bn.g1_add(acc, None if (inf is not None and inf[i]) else L.g1_from(pts_xy[i]))
        return L.g1_m(acc), 1 if acc is None else 0


def _worker(rank, world, port, n, seed, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(seed)
    tau = rng.randrange(1, bn.R)
    srs = [bn.g1_mul(bn.G1_GEN, pow(tau, i, bn.R)) for i in range(n)]
    scalars = [rng.randrange(bn.R) for _ in range(n)]
    scalars[0] = 0
    lo, hi = kd.shard_range(n, rank, world)
    local = L.fr_vec(scalars[lo:hi]).reshape(hi - lo, 8)
    xy, inf = kd.sharded_commit(OracleCtx(srs), local, lo)
    want = bn.g1_mul(bn.G1_GEN, sum(s * pow(tau, i, bn.R) for i, s in enumerate(scalars)) % bn.R)
    out[rank] = (None if inf else L.g1_from(xy)) == want
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("n", [7, 10])
def test_sharded_commit_world2_gloo(n):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, 1234 + n, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_ranges_are_a_partition():
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) - 1):
        for world in (1, 2, 3, 4, 8):
            r = [kd.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_partial_packing_roundtrip():
    xy = np.arange(16, dtype=np.uint32) * 0x10203041
    buf = np.concatenate([kd.pack_partial(xy, 0), kd.pack_partial(np.zeros(16, np.uint32), 1)])
    pts, infs = kd.unpack_partials(buf)
    assert np.array_equal(pts[0], xy) and infs.tolist() == [0, 1]
