"""N>1 host logic on CPU: world_size-2 gloo.  Columns shard with no data-path collective; per-column inputs do
not depend on the sharding; the only collectives are the MAX of timings and the ensemble diagnostics reduction."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_perturbations_are_shard_invariant():
    sys.path.insert(0, str(ROOT))
    import bench
    site_a, sc_a, of_a, amp_a = bench.perturbations(0, 1024)
    site_b, sc_b, of_b, amp_b = bench.perturbations(512, 512)
    assert np.array_equal(site_a[512:], site_b) and np.array_equal(sc_a[:, 512:], sc_b)
    assert np.array_equal(of_a[:, 512:], of_b) and np.array_equal(amp_a[512:], amp_b)
    assert set(np.unique(site_a)) == set(range(9))
    assert (sc_a[2] == 1).all() and (of_a[[0, 1, 3]] == 0).all()   # T2m is offset-only, the fluxes scale-only
    # weak scaling: rank r owns global columns [r*TOTAL, (r+1)*TOTAL); block 0 is the config-5 ensemble and a range
    # that crosses a block boundary is the concatenation of the two blocks' numbers
    T = bench.TOTAL_COLUMNS
    s_x, sc_x, of_x, amp_x = bench.perturbations(T - 3, 8)
    s_0, sc_0, of_0, amp_0 = bench.perturbations(T - 3, 3)
    s_1, sc_1, of_1, amp_1 = bench.perturbations(T, 5)
    assert np.array_equal(np.concatenate([amp_0, amp_1]), amp_x) and np.array_equal(np.concatenate([sc_0, sc_1], 1), sc_x)
    assert np.array_equal(np.concatenate([of_0, of_1], 1), of_x) and np.array_equal(np.concatenate([s_0, s_1]), s_x)
    assert not np.array_equal(bench.perturbations(0, 64)[3], amp_1[:5].repeat(13)[:64])  # block 1 is not block 0 again
    assert not np.array_equal(bench.perturbations(0, 5)[3], amp_1)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np, torch, torch.distributed as dist
    import bench
    from oracle import oracle
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    total = 8
    per = total // world
    col0 = rank * per
    st = bench.load_state(bench.START_RECORD)
    sites = bench.load_sites(64)
    # each rank advances ITS columns (here with the CPU oracle standing in for the engine)
    cols = []
    for c in range(col0, col0 + per):
        s, sc, of, am = bench.perturbations(c, 1)
        col = oracle.Column(4, "det")
        col.set_forcing(*[sites[s[0], k] * sc[k, 0] + of[k, 0] for k in range(4)])
        col.load_state(st); col.set_scalar("oflux_amp", float(am[0])); col.step(50)
        cols.append(col)
    th = torch.tensor([c.scalar("thickness") for c in cols], dtype=torch.float64)
    # the optional diagnostics gather / reduction is the only collective
    gathered = [torch.zeros(per, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, th)
    red = torch.tensor([th.sum().item(), -th.min().item(), th.max().item()], dtype=torch.float64)
    s = red[:1].clone(); dist.all_reduce(s, op=dist.ReduceOp.SUM)
    mm = red[1:].clone(); dist.all_reduce(mm, op=dist.ReduceOp.MAX)
    t = torch.tensor([0.5 + rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(%(out)r, np.concatenate([torch.cat(gathered).numpy(), [s.item(), -mm[0].item(), mm[1].item(), t.item()]]))
    dist.destroy_process_group()
""")


def test_two_rank_gloo_sharding_matches_single_process(tmp_path, oracle_mod):
    sys.path.insert(0, str(ROOT))
    import bench
    out = tmp_path / "res.npy"
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": str(ROOT), "out": str(out)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    res = np.load(out)
    # single-process reference: the same 8 columns
    st = bench.load_state(bench.START_RECORD)
    sites = bench.load_sites(64)
    th = []
    for c in range(8):
        s, sc, of, am = bench.perturbations(c, 1)
        col = oracle_mod.Column(4, "det")
        col.set_forcing(*[sites[s[0], k] * sc[k, 0] + of[k, 0] for k in range(4)])
        col.load_state(st); col.set_scalar("oflux_amp", float(am[0])); col.step(50)
        th.append(col.scalar("thickness"))
    th = np.array(th)
    assert np.array_equal(res[:8], th)                 # shards give identical per-column results
    assert res[8] == th[:4].sum() + th[4:].sum() and res[9] == th.min() and res[10] == th.max()
    assert res[11] == 1.5                              # MAX over ranks of the timing


def test_shard_partition_covers_all_columns():
    from samsim_b200 import distributed as D
    for total, world in ((1 << 20, 8), (10, 4), (7, 8), (100, 3)):
        spans = [D.shard(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(n for _, n in spans) == total
        for (a, n), (b, _) in zip(spans, spans[1:]):
            assert a + n == b


WORKER2 = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np, torch, torch.distributed as dist
    from samsim_b200 import distributed as D
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    col0, n = D.shard(11, rank, world)
    vals = np.arange(col0, col0 + n, dtype=np.float64) + 1.0
    local = {k: {"sum": float(vals.sum()), "min": float(vals.min()), "max": float(vals.max())} for k in D.NAMES}
    red = D.reduce_ensemble(local, n)
    g = D.gather_columns(torch.from_numpy(vals))
    if rank == 0:
        np.save(%(out)r, np.array([red["thickness"]["mean"], red["thickness"]["min"], red["thickness"]["max"], red["columns"]] + g.tolist()))
    dist.destroy_process_group()
""")


def test_ensemble_reduction_and_gather_two_ranks(tmp_path):
    out = tmp_path / "r.npy"
    script = tmp_path / "w.py"
    script.write_text(WORKER2 % {"root": str(ROOT), "out": str(out)})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29542", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=120) == 0
    r = np.load(out)
    assert r[0] == 6.0 and r[1] == 1.0 and r[2] == 11.0 and r[3] == 11
    assert np.array_equal(r[4:], np.arange(1, 12, dtype=np.float64))   # uneven shards (6 + 5) gathered in order
