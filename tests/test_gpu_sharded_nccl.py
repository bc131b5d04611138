"""One region's events split across GPUs, the per-mutation sums combined over NCCL INSIDE the library
(ps_score_mutations_sharded, csrc/ps_comm.cu; cpp/MakeMutations.cpp:19-22, 38-52; poreseq/Variant.py:71-76).

One process per GPU (spawned here), the NCCL id travels through a file.  ordered mode: bit-identical to the single-GPU
scores (and so to the reference, which the single-GPU path is pinned to); all-reduce mode: ~1e-16 relative.  Both
precision modes.  With one visible GPU the same calls run on a one-rank communicator."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, id_path, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import time
    from poreseq_b200 import poreseqcpp, sharding, synth
    from util import edge_mutations
    reg = synth.make_region(2000, 12, seed=21, draft_error=0.02, partial=0.3, params=dict(scoring_width=100))
    rng = np.random.default_rng(77)
    st, og, mu = synth.random_mutations(reg.sequence, 400, rng, max_len=4)
    e_st, e_og, e_mu = edge_mutations(reg.sequence, 5, count=0)
    st, og, mu = st + e_st, og + e_og, mu + e_mu
    res = {}
    for ordered in (True, False):
        if rank == 0:
            uid = poreseqcpp.comm_unique_id()
            with open(id_path + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(id_path + ".tmp", id_path + (".o" if ordered else ".a"))
        else:
            p = id_path + (".o" if ordered else ".a")
            for _ in range(600):
                if os.path.exists(p):
                    break
                time.sleep(0.05)
            uid = open(p, "rb").read()
        ctx = poreseqcpp.Context(rank)
        ctx.comm_init(uid, rank, world, ordered=ordered)
        for precision in ("exact", "fast"):
            ctx.set_precision(precision)
            res[(ordered, precision)] = sharding.score_mutations_sharded(ctx, reg, st, og, mu, rank, world)
        ctx.comm_destroy()
        ctx.close()
    if rank == 0:
        single = {}
        for precision in ("exact", "fast"):
            c = poreseqcpp.Context(0)
            c.set_precision(precision)
            nr = poreseqcpp.NativeRegion(c, reg.sequence, reg.events, reg.params)
            single[precision] = nr.score_mutations(st, og, mu)
            nr.close(); c.close()
        np.savez(out_path, **{"%s_%s" % ("ord" if o else "all", p): v for (o, p), v in res.items()},
                 single_exact=single["exact"], single_fast=single["fast"])
    else:
        np.savez(out_path + ".r%d.npz" % rank, **{"%s_%s" % ("ord" if o else "all", p): v for (o, p), v in res.items()})


def _run(world, tmp_path):
    id_path, out_path = str(tmp_path / "nccl_id"), str(tmp_path / "out.npz")
    ctxm = mp.get_context("spawn")
    procs = [ctxm.Process(target=_worker, args=(r, world, id_path, out_path)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    z = np.load(out_path)
    ex, fa = z["single_exact"], z["single_fast"]
    # ordered: the sum over events runs in event order across the ranks -> the same bits as one GPU
    assert np.array_equal(z["ord_exact"], ex)
    assert np.array_equal(z["ord_fast"], fa)
    # all-reduce of partial sums: different association of the FP64 sum
    assert np.allclose(z["all_exact"], ex, rtol=1e-12, atol=1e-12) and np.array_equal(z["all_exact"] >= 0, ex >= 0)
    assert np.allclose(z["all_fast"], fa, rtol=1e-12, atol=1e-9) and np.array_equal(z["all_fast"] >= 0, fa >= 0)
    for r in range(1, world):
        zr = np.load(out_path + ".r%d.npz" % r)
        for k in ("ord_exact", "ord_fast", "all_exact", "all_fast"):
            assert np.array_equal(zr[k], z[k]), "rank %d disagrees with rank 0 on %s" % (r, k)


def _gpus():
    import ctypes
    try:
        cuda = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.device_count()
        except Exception:
            return 1
    n = ctypes.c_int(0)
    cuda.cudaGetDeviceCount(ctypes.byref(n))
    return n.value


def test_sharded_one_rank(tmp_path):
    _run(1, tmp_path)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_nccl(world, tmp_path):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    _run(world, tmp_path)
