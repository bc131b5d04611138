"""The randomised degenerate-region sweep (scripts/gpu_sweep.py) as part of the GPU suite: 300 seeded tiny regions
(events without alignment, with a single level or a single aligned level, non-ACGT bases, bands of a few rows) through
ScoreAlignments, ScorePoints, ScoreMutations and Refine in both precisions against the CPU checker."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scripts"))


def test_random_degenerate_regions_match_the_checker(capsys):
    import gpu_sweep
    rc = gpu_sweep.main(n=300, first=0)
    out = capsys.readouterr().out
    assert rc == 0, out[-4000:]


def test_random_degenerate_regions_driver_entry_points(capsys):
    """swfull on the device, MapAlignments, FindMutations (repeated seed), the Mutate loop and ViterbiMutate (best path and
    8 sampled walks on one rand() stream) on 120 seeded degenerate regions, incl. transition probabilities above 1."""
    import gpu_sweep
    rc = gpu_sweep.main_drivers(n=120, first=0)
    out = capsys.readouterr().out
    assert rc == 0, out[-4000:]


def test_random_mid_sized_regions_batched_entry_points(capsys):
    """150-500 bases, realign_width 20-300, jittered / partial / missing alignments: ScorePoints through the batched and
    the handle-less entry points, ScoreEvents (FP32 score-only fill in fast mode), Refine."""
    import gpu_sweep
    rc = gpu_sweep.main_mid(n=12, first=0)
    out = capsys.readouterr().out
    assert rc == 0, out[-4000:]


def test_random_small_regions_whole_consensus_loop(capsys):
    """ps_consensus_batch (lockstep over regions, region-private rand() streams) against the Mutate.py policy driven
    through the checker, 8 small regions in both precisions."""
    import gpu_sweep
    rc = gpu_sweep.main_consensus(n=8, first=0)
    out = capsys.readouterr().out
    assert rc == 0, out[-4000:]


def test_random_regions_through_the_psalign_mirror(capsys):
    """PSAlign.ScoreMutations -> ApplyMuts (MakeMutations' recursion on a caller-scored list of multi-base edits) and Copy
    on 60 tiny regions, against the checker."""
    import gpu_sweep
    rc = gpu_sweep.main_psalign(n=60, first=0)
    out = capsys.readouterr().out
    assert rc == 0, out[-4000:]
