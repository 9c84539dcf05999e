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


def _shard_worker(rank, world, port, emu_path, out_dir, low_complexity):
    """Sharded build (every rank sorts its slice of the k-mer ranks, postings all-gathered) vs one single-process index."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    from pandelos_b200 import digest, multigpu, native, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    native.load(emu_path)
    w = synth.generate(6, 25, 60.0, 0.1, 93, low_complexity=low_complexity)
    k = 3
    data = native.PangeneIData(w.residues, w.offsets, w.genome_of)
    if low_complexity > 0:   # the input "resident on the device" (host memory under the emulator): residues and the gene table by pointer
        pn, bounds = multigpu.build_sharded(dist, native, k, data, residues_device_ptr=data.residues.ctypes.data,
                                            table_device_ptrs=(data.offsets.ctypes.data, data.sequenceGenome.ctypes.data))
    else:
        pn, bounds = multigpu.build_sharded(dist, native, k, data)
    one = native.PangeneNative(k, data)
    ok = pn.info.U == one.info.U and pn.info.lookups == one.info.lookups and pn.info.groups == one.info.groups and pn.info.N == one.info.N
    gb = multigpu.genome_bounds(w.genome_of, one.info.G)
    ok = ok and set(bounds.tolist()) <= set(gb.tolist()) and bounds[0] == 0 and bounds[-1] == one.info.S
    kl, tv = pn.gene_stats()
    kl1, tv1 = one.gene_stats()
    ok = ok and (kl == kl1).all() and (tv == tv1).all()
    g0, g1 = int(np.searchsorted(gb, bounds[rank])), int(np.searchsorted(gb, bounds[rank + 1]))
    cells = 0
    for g in range(one.info.G):
        if g0 <= g < g1:
            a, b = pn.generateScoresPart(g), one.generateScoresPart(g)
            ok = ok and digest.scores_digest(a) == digest.scores_digest(b) and pn.last_stats.pairs == one.last_stats.pairs
            src, dst, sc, _ = pn.genomeEdges(g)
            src1, dst1, sc1, _ = one.genomeEdges(g)
            ok = ok and sorted(zip(src.tolist(), dst.tolist(), sc.view(np.uint32).tolist())) == sorted(zip(src1.tolist(), dst1.tolist(), sc1.view(np.uint32).tolist()))
            cells += a.scoresCount
        else:
            try:
                pn.generateScoresPart(g)
                ok = False
            except native.PdError as e:
                ok = ok and e.code == native.PD_ERR_INVALID
    if g1 > g0:
        st = pn.score_partition_device(int(bounds[rank]), int(bounds[rank + 1]))
        ok = ok and st.cells == cells
    open(os.path.join(out_dir, "result%d" % rank), "w").write("ok %d" % cells if ok else "mismatch")
    pn.close()
    one.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,low", [(2, 0.0), (3, 0.3), (4, 0.1), (8, 0.0)])
def test_sharded_build_matches_the_single_process_index(tmp_path, world, low):
    import build_emu
    import torch.multiprocessing as mp
    emu = build_emu.build()
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_shard_worker, args=(world, port, emu, str(tmp_path), low), nprocs=world, join=True)
    res = [open(os.path.join(str(tmp_path), "result%d" % r)).read() for r in range(world)]
    assert all(r.startswith("ok") for r in res), res
    assert sum(int(r.split()[1]) for r in res) > 0


def _tiny_worker(rank, world, port, emu_path, out_dir):
    """An input too small for the rank count: every rank must refuse, at the same point, with an error — not hang."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    from pandelos_b200 import multigpu, native, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    native.load(emu_path)
    w = synth.from_sequences(["AAAAAA", "AAAAAA"], [0, 1])   # every k-mer is the same: one slice gets everything
    try:
        multigpu.build_sharded(dist, native, 3, native.PangeneIData(w.residues, w.offsets, w.genome_of))
        res = "built"
    except native.PdError as e:
        res = "refused" if e.code == native.PD_ERR_UNSUPPORTED else "error %d" % e.code
    open(os.path.join(out_dir, "tiny%d" % rank), "w").write(res)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_build_refuses_too_small_inputs_on_every_rank(tmp_path):
    import build_emu
    import torch.multiprocessing as mp
    emu = build_emu.build()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_tiny_worker, args=(2, port, emu, str(tmp_path)), nprocs=2, join=True)
    assert [open(os.path.join(str(tmp_path), "tiny%d" % r)).read() for r in range(2)] == ["refused", "refused"]
