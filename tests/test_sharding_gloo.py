"""world_size-2 gloo test (CPU) of the event-shard all-reduce and the region assignment.
The per-shard scorer here is the CPU checker (test infrastructure); on the GPU the same host logic
runs with NativeRegion.score_mutations_partial (tests/test_gpu_parity.py::test_event_sharded_partials)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from oracle import binding
    from poreseq_b200 import sharding, synth
    from util import SMALL
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = binding.load("oracle")
    reg = synth.make_region(length=300, coverage=3, seed=5, draft_error=0.03, params=SMALL)
    st, og, mu = synth.point_mutations(reg.sequence)
    st, og, mu = st[:400], og[:400], mu[:400]

    def partial(shard, s, o, m):
        sc, _ = orc.score_mutations(shard, s, o, m)
        return sc + 1e-6          # the checker starts its sums at -1e-6

    got = sharding.score_mutations_event_sharded(reg, st, og, mu, partial, rank, world)
    want, _ = orc.score_mutations(reg, st, og, mu)
    ok = np.allclose(got, want, rtol=1e-9, atol=1e-9) and np.array_equal(got >= 0, want >= 0)
    regions = sharding.assign_regions(7, rank, world)
    out[rank] = (bool(ok), regions, sharding.event_block(len(reg.events), rank, world))
    dist.destroy_process_group()


def test_event_shard_allreduce_world2():
    from oracle import binding
    binding.build("oracle")
    world, port = 2, 29517 + os.getpid() % 500
    with mp.get_context("spawn").Manager() as mgr:      # (the test process may already own worker threads: no fork)
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] and res[1][0]
    assert sorted(res[0][1] + res[1][1]) == list(range(7))
    assert res[0][2] == (0, 3) and res[1][2] == (3, 6)
