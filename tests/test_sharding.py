"""Length-balanced utterance sharding (the only multi-GPU logic of this path: no
collective) incl. a world_size-2 gloo run of the rank-side logic.  CPU only."""
import importlib
import os
import sys

import numpy as np
import pytest


def test_lpt_is_a_partition_and_balanced(synth):
    sh = importlib.import_module("speaker-embedding-with-phonetic-information_b200.shard")
    lengths = synth.make_lengths(10000, 1000, 30000, 1238)
    for n in (1, 2, 4, 8):
        shards = sh.lpt_shards(lengths, n)
        allidx = np.concatenate(shards)
        assert sorted(allidx.tolist()) == list(range(len(lengths)))
        for s in shards:
            assert np.all(np.diff(s) > 0)                    # input order inside a shard
        loads = sh.shard_loads(lengths, shards)
        assert loads.max() - loads.min() <= lengths.max()
        assert loads.max() / loads.mean() < 1.001            # near-perfect balance at this size


def test_lpt_edge_cases():
    sh = importlib.import_module("speaker-embedding-with-phonetic-information_b200.shard")
    assert [s.tolist() for s in sh.lpt_shards([], 3)] == [[], [], []]
    assert [s.tolist() for s in sh.lpt_shards([5], 2)] == [[0], []]
    with pytest.raises(ValueError):
        sh.lpt_shards([1, 2], 0)
    shards = sh.lpt_shards([3, 3, 3, 3], 2)
    parts = [np.full((len(s), 2), i, np.float32) for i, s in enumerate(shards)]
    merged = sh.merge_in_input_order(4, shards, parts)
    assert merged.shape == (4, 2) and set(merged[:, 0].tolist()) == {0.0, 1.0}


def _rank_main(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sh = importlib.import_module("speaker-embedding-with-phonetic-information_b200.shard")
    sy = importlib.import_module("speaker-embedding-with-phonetic-information_b200.synth")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = sy.make_lengths(world * 50, 1000, 30000, 1238)
    mine = sh.lpt_shards(lengths, world)[rank]
    # what bench.py reduces: per-rank audio seconds (SUM) and per-rank time (MAX)
    audio = torch.tensor([lengths[mine].sum() / 100.0], dtype=torch.float64)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(audio, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # per-rank results come back in input order without any data-path collective: write + merge
    np.save(os.path.join(tmp, f"part{rank}.npy"), np.stack([mine.astype(np.float32), np.full(len(mine), rank, np.float32)], 1))
    dist.barrier()
    if rank == 0:
        shards = sh.lpt_shards(lengths, world)
        parts = [np.load(os.path.join(tmp, f"part{r}.npy")) for r in range(world)]
        merged = sh.merge_in_input_order(len(lengths), shards, parts)
        assert np.array_equal(merged[:, 0], np.arange(len(lengths), dtype=np.float32))
        assert abs(audio.item() - lengths.sum() / 100.0) < 1e-6 and t.item() == float(world)
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_rank_main, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
