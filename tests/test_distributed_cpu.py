"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: range split and the variable-length all-gather."""

import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_molecule_range_partitions():
    from nvmolkit_b200.distributed import molecule_range

    for n in (0, 1, 7, 8, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [molecule_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import torch, torch.distributed as dist
    from nvmolkit_b200.distributed import all_gather_v, molecule_range, rank_world
    dist.init_process_group("gloo")
    rank, world = rank_world()
    assert world == 2
    # each rank owns a molecule range and produces a variable number of result rows per molecule
    lo, hi = molecule_range(11, rank, world)
    rows = torch.tensor([[m, k] for m in range(lo, hi) for k in range(m % 3 + 1)], dtype=torch.int32).reshape(-1, 2)
    allrows, sizes = all_gather_v(rows)
    want = torch.tensor([[m, k] for m in range(11) for k in range(m % 3 + 1)], dtype=torch.int32)
    assert torch.equal(allrows, want), (allrows, want)
    assert int(sizes.sum()) == want.shape[0]
    # empty contribution from one rank
    part = torch.arange(5, dtype=torch.float64).reshape(5, 1) if rank == 1 else torch.empty((0, 1), dtype=torch.float64)
    got, _ = all_gather_v(part)
    assert torch.equal(got, torch.arange(5, dtype=torch.float64).reshape(5, 1))
    counts = torch.full((4,), rank + 1, dtype=torch.int32)
    dist.all_reduce(counts)
    assert counts.tolist() == [3, 3, 3, 3]
    from nvmolkit_b200.distributed import map_molecule_range
    rows = map_molecule_range(11, lambda lo, hi: torch.arange(lo, hi, dtype=torch.int64).reshape(-1, 1) * 10)
    assert rows.ravel().tolist() == [10 * i for i in range(11)]  # item order preserved, every rank has everything
    rows = map_molecule_range(1, lambda lo, hi: torch.arange(lo, hi, dtype=torch.int64).reshape(-1, 1))  # one rank gets nothing
    assert rows.ravel().tolist() == [0]
    from nvmolkit_b200.distributed import sharded_upload
    for n in (1, 7, 8, 13):
        h = torch.arange(n * 3, dtype=torch.int32).reshape(n, 3)
        assert torch.equal(sharded_upload(h, "cpu"), h)  # every rank ends up with all rows although it copied only its slice
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_all_gather_v_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("ok") == 2
