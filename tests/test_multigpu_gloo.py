"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path — volume-balanced genome-snapped row
partition, per-rank scoring of its partition into a best-hit slice, all-gather of the slices — with the kernels'
logic supplied by the SIMT emulator build (tests/emu).  The gathered table must equal the best-hit rows of
single-process `computeScores` calls, bit for bit."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))


def _worker(rank, world, port, emu_path, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    from pandelos_b200 import multigpu, native, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    native.load(emu_path)
    w = synth.generate(5, 30, 70.0, 0.1, 91)
    k = 3
    pn = native.PangeneNative(k, native.PangeneIData(w.residues, w.offsets, w.genome_of))
    S, G = pn.info.S, pn.info.G
    gb = multigpu.genome_bounds(w.genome_of, G)
    _, visited = pn.gene_stats()
    bounds = multigpu.balanced_bounds(visited, 0, S, world, snap=gb)
    rows = [int(bounds[r + 1] - bounds[r]) for r in range(world)]
    bh_local = torch.zeros((max(rows), G), dtype=torch.float32)
    st = pn.score_partition_device(int(bounds[rank]), int(bounds[rank + 1]), best_hit_ptr=bh_local.data_ptr())
    bh_all, _ = multigpu.allgather_best_hits(dist, bh_local, rows, G, "cpu")
    pairs = torch.tensor([float(st.pairs)], dtype=torch.float64)
    dist.all_reduce(pairs)
    # the overlapped variant: scoring and gathering chunk by chunk (3 chunks, the last one ragged) must give the same table
    gather = multigpu.ChunkedBestHitGather(dist, rows, G, "cpu", chunks=3)
    tot = multigpu.score_and_gather(pn, gather, rank, int(bounds[rank]))
    chunked = gather.assemble()
    same = torch.equal(chunked.view(torch.int32), bh_all.view(torch.int32)) and tot["pairs"] == st.pairs and tot["rows"] == rows[rank]
    probe_rank, probe_row = world - 1, rows[world - 1] - 1
    same = same and torch.equal(gather.table.view(-1, G)[gather.offset_of(probe_rank, probe_row)],
                                bh_all[sum(rows[:probe_rank]) + probe_row])
    flag = torch.tensor([1.0 if same else 0.0], dtype=torch.float64)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        want = np.zeros((S, G), np.float32)
        total = 0
        for g in range(G):
            s = pn.generateScoresPart(g)
            want[np.nonzero(w.genome_of == g)[0]] = s.max_genome_score
            total += pn.last_stats.pairs
        got = bh_all.numpy()
        ok = got.shape == want.shape and (got.view(np.uint32) == want.view(np.uint32)).all() and int(pairs.item()) == total \
            and set(bounds.tolist()) <= set(gb.tolist()) and (want > 0).any() and flag.item() == 1.0
        open(os.path.join(out_dir, "result"), "w").write("ok" if ok else "mismatch")
    pn.close()
    dist.barrier()
    dist.destroy_process_group()


def test_partition_score_allgather_world2(tmp_path):
    import build_emu
    import torch.multiprocessing as mp
    emu = build_emu.build()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, emu, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "result")).read() == "ok"


def test_balanced_bounds_properties():
    from pandelos_b200 import multigpu
    rng = np.random.default_rng(5)
    visited = rng.integers(0, 1000, size=5000).astype(np.uint64)
    genome_of = np.repeat(np.arange(50), 100)
    gb = multigpu.genome_bounds(genome_of, 50)
    for parts in (1, 2, 4, 8):
        b = multigpu.balanced_bounds(visited, 0, 5000, parts)
        assert b[0] == 0 and b[-1] == 5000 and (np.diff(b) >= 0).all()
        vol = np.add.reduceat(visited.astype(np.float64) + 1, b[:-1])
        assert vol.max() / vol.mean() < 1.05
        bs = multigpu.balanced_bounds(visited, 0, 5000, parts, snap=gb)
        assert set(bs.tolist()) <= set(gb.tolist())
    with pytest.raises(ValueError):
        multigpu.genome_bounds(np.array([0, 1, 0]), 2)
