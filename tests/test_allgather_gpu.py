"""b200mol_allgather_counts / b200mol_allgather_results: the NCCL result exchange of the C-ABI, driven without torch's
collectives (the communicator is created straight on NCCL through ctypes, as a C++ caller would). World size 1 runs on the
single-GPU test box; world size 2 runs when two GPUs are visible (gpurun --gpus 2)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _nccl():
    lib = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)  # the copy torch already loaded (same soname), else the system one
    lib.ncclGetUniqueId.argtypes = [C.POINTER(_UniqueId)]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    return lib


def _exchange(rank, world, uid_bytes, out_queue=None):
    from nvmolkit_b200 import _lib

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    nccl = _nccl()
    uid = _UniqueId()
    C.memmove(C.byref(uid), uid_bytes, 128)
    comm = C.c_void_p()
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    rng = np.random.default_rng(100 + rank)
    n_conf = 3 + 2 * rank
    atoms = rng.integers(4, 30, n_conf).astype(np.int32)
    pos = rng.normal(0, 1, (int(atoms.sum()), 3))
    en = rng.normal(0, 1, n_conf)
    conv = (rng.random(n_conf) < 0.5).astype(np.int8)
    d_pos, d_atoms = torch.from_numpy(pos).to(dev), torch.from_numpy(atoms).to(dev)
    d_en, d_conv = torch.from_numpy(en).to(dev), torch.from_numpy(conv).to(dev)
    s = torch.cuda.current_stream().cuda_stream
    cc, ac = (C.c_int64 * world)(), (C.c_int64 * world)()
    _lib.call("b200mol_allgather_counts", comm, n_conf, int(atoms.sum()), cc, ac, s)
    tot_c, tot_a = sum(cc), sum(ac)
    o_pos = torch.empty((tot_a, 3), dtype=torch.float64, device=dev)
    o_atoms = torch.empty(tot_c, dtype=torch.int32, device=dev)
    o_en = torch.empty(tot_c, dtype=torch.float64, device=dev)
    o_conv = torch.empty(tot_c, dtype=torch.int8, device=dev)
    _lib.call("b200mol_allgather_results", comm, cc, ac, d_pos.data_ptr(), d_atoms.data_ptr(), d_en.data_ptr(), d_conv.data_ptr(),
              o_pos.data_ptr(), o_atoms.data_ptr(), o_en.data_ptr(), o_conv.data_ptr(), s)
    torch.cuda.synchronize()
    # what every rank must hold: the rank-major concatenation of what each rank generated from its seed
    want_pos, want_atoms, want_en, want_conv = [], [], [], []
    for r in range(world):
        g = np.random.default_rng(100 + r)
        nc = 3 + 2 * r
        a = g.integers(4, 30, nc).astype(np.int32)
        want_atoms.append(a)
        want_pos.append(g.normal(0, 1, (int(a.sum()), 3)))
        want_en.append(g.normal(0, 1, nc))
        want_conv.append((g.random(nc) < 0.5).astype(np.int8))
    ok = (list(cc) == [3 + 2 * r for r in range(world)] and np.array_equal(o_atoms.cpu().numpy(), np.concatenate(want_atoms))
          and np.array_equal(o_pos.cpu().numpy(), np.concatenate(want_pos)) and np.array_equal(o_en.cpu().numpy(), np.concatenate(want_en))
          and np.array_equal(o_conv.cpu().numpy(), np.concatenate(want_conv)))
    nccl.ncclCommDestroy(comm)
    if out_queue is not None:
        out_queue.put((rank, ok))
    return ok


def test_allgather_results_single_rank(cuda):
    uid = _UniqueId()
    assert _nccl().ncclGetUniqueId(C.byref(uid)) == 0
    assert _exchange(0, 1, C.string_at(C.byref(uid), 128))


def _worker(rank, world, uid_bytes, q):
    _exchange(rank, world, uid_bytes, q)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allgather_results_two_ranks(cuda):
    import torch.multiprocessing as mp

    uid = _UniqueId()
    assert _nccl().ncclGetUniqueId(C.byref(uid)) == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, C.string_at(C.byref(uid), 128), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]
