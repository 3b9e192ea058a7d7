#!/usr/bin/env python
"""bench.py — BASELINE.json metric on the BASELINE.json configuration.

Default workload (configs[1]): 1M x 1M symmetric 2048-bit Tanimoto + Butina clustering (similarity >= 0.7, i.e.
distance cutoff 0.3), synthetic clustered fingerprints (20,000 centres x 50 members, seed 20260924), one GPU.
A step = one pass of the hot path over the whole batch: thresholded similarity graph (counts + edge list, every
unordered pair evaluated once) -> CSR -> greedy Butina loop -> cluster ids. `value` = unique pairs / s with the
fingerprints resident in HBM; `e2e` = the same through the public API with HOST buffers (pinned H2D of the fingerprints
and D2H of the ids inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload butina|etkdg_mmff]
    torchrun --nproc-per-node N bench.py --gpus N ...     (one rank per GPU; rank 0 prints the JSON line)
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CUTOFF = 0.3  # similarity threshold 0.7
METRIC_NAME = "tanimoto_pairs_per_s"
UNIT = "pairs/s"


def unique_pairs(n: int) -> float:
    return n * (n - 1) / 2.0


# ----------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_sample_size(words: int = 64, target_s: float = 10.0) -> tuple[int, float]:
    """Calibrate the oracle on a small set, then size a sample worth ~target_s of CPU work (multiple of 50 rows)."""
    import oracle
    from nvmolkit_b200 import synthetic

    fp = synthetic.clustered_fingerprints(160, 50, seed=synthetic.SEED + 1)
    t0 = time.perf_counter()
    oracle.butina_fp(fp, CUTOFF)
    rate = unique_pairs(len(fp)) / (time.perf_counter() - t0)
    n = int((2.0 * rate * target_s) ** 0.5)
    n = max(2000, min(100_000, n // 50 * 50))
    return n, rate


def run_reference(args) -> None:
    """The reference's CPU implementation of the path, timed on this box's host cores.

    The reference's path is RDKit (BulkTanimotoSimilarity + ML.Cluster.Butina.ClusterData(reordering=True)); RDKit
    cannot be built or imported here, so the arm times the C restatement in oracle/ ("port") with OpenMP on all cores,
    on a bounded sample of the same workload.
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from nvmolkit_b200 import synthetic

    cores = oracle.set_threads(os.cpu_count() or 1)  # (torchrun exports OMP_NUM_THREADS=1 to its workers)
    n, _ = cpu_sample_size()
    if args.n_centres:
        n = min(n, args.n_centres * 50)
    fp = synthetic.clustered_fingerprints(n // 50, 50, seed=synthetic.SEED)
    for _ in range(args.warmup):
        oracle.butina_fp(fp, CUTOFF)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ids, cen = oracle.butina_fp(fp, CUTOFF)
    dt = (time.perf_counter() - t0) / args.steps
    value = unique_pairs(len(fp)) / dt
    sample = f"{len(fp)} x {len(fp)} clustered 2048-bit fingerprints ({len(fp) // 50} centres x 50), cutoff {CUTOFF}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32 popcount + f64 divide", "data": "synthetic",
        "config": {"workload": "1Mx1M symmetric 2048-bit Tanimoto + Butina (sim>=0.7), CPU on a bounded sample",
                   "sample_rows": len(fp), "cutoff": CUTOFF},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "n_clusters": int(len(cen)),
    }))


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    from nvmolkit_b200 import _lib, synthetic
    from nvmolkit_b200.clustering import fused_butina_device, fused_butina_sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: nvmolkit_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    _lib.check(_lib.load().b200mol_check_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    n_centres = args.n_centres or 20000
    fp_host = synthetic.clustered_fingerprints(n_centres, 50, seed=synthetic.SEED)
    n, words = fp_host.shape
    h_fp = torch.from_numpy(fp_host.view(np.int32)).pin_memory()
    d_fp = h_fp.to(dev)
    h_ids = torch.empty(n, dtype=torch.int32).pin_memory()
    stream = torch.cuda.current_stream()

    def step_device(x):
        if world == 1:
            return fused_butina_device(x, CUTOFF)
        return fused_butina_sharded(x, CUTOFF)

    def step_e2e():
        x = h_fp.to(dev, non_blocking=True)
        ids, cen = step_device(x)
        h_ids.copy_(ids, non_blocking=True)
        stream.synchronize()
        return ids, cen

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps, out

    _lib.profile_enable(True)
    for _ in range(args.warmup):
        step_device(d_fp)
    launches0 = _lib.launch_count()
    with ClockSampler(local) as clocks:
        ms_dev, (ids, cen) = timed(lambda: step_device(d_fp), args.steps)
        launches = _lib.launch_count() - launches0
        # dominant kernel alone (CUDA events on its own stream, recorded inside the library around the tile kernel)
        pass_ms = []
        tensor_path = True
        for _ in range(max(1, min(3, args.steps))):
            step_device(d_fp)
            try:
                pass_ms.append(_lib.profile_read("neighbor_pass_tc"))
            except ValueError:
                tensor_path = False
                pass_ms.append(_lib.profile_read("neighbor_pass"))
        phases = {k: _lib.profile_read(k) for k in ("neighbor_pass", "csr_build", "cluster_loop")}
        if tensor_path:
            phases["neighbor_pass_tc"] = _lib.profile_read("neighbor_pass_tc")
        step_e2e()
        ms_e2e, _ = timed(step_e2e, args.steps)

    # materialised cross-similarity (the reference's crossTanimotoSimilarity output format): HBM-write bound, 8 B / pair
    cross = None
    if args.cross_n > 0 and rank == 0:
        from nvmolkit_b200.similarity import crossTanimotoSimilarity

        xa = d_fp[: args.cross_n]
        xb = d_fp[args.cross_n: 2 * args.cross_n] if n >= 2 * args.cross_n else d_fp[: args.cross_n]
        for _ in range(2):
            res = crossTanimotoSimilarity(xa, xb)
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            res = crossTanimotoSimilarity(xa, xb)
            try:
                times.append(_lib.profile_read("cross_tc"))
            except ValueError:
                times = []
                break
        if times:
            nb = xb.shape[0]
            ms_c = float(np.mean(times))
            bytes_c = 8.0 * args.cross_n * nb + 256.0 * (args.cross_n + nb)
            peak_c, src_c = measured_peaks()
            cross = {"pairs_per_s": args.cross_n * nb / (ms_c * 1e-3), "kernel_ms": ms_c, "shape": [args.cross_n, nb],
                     "roofline": {"bound": "hbm", "achieved": bytes_c / (ms_c * 1e-3) / 1e9, "peak": peak_c, "unit": "GB/s",
                                  "frac": bytes_c / (ms_c * 1e-3) / 1e9 / peak_c, "algorithmic_bytes": bytes_c,
                                  "kernel": "simTensorKernel<materialise> (cross_tc)", "peak_source": src_c}}
        del res

    # second half of the BASELINE metric: ETKDG + MMFF mols/s (config 3 shape, reduced count so the default run stays short)
    path_b = None
    if args.etkdg_mols > 0:
        flat_b, mmff_b = path_b_pool(args.pool, synthetic.SEED)
        path_b = run_path_b_gpu(flat_b, mmff_b, args.etkdg_mols, args.confs, dev, max(1, args.steps - 1), 2, world, rank)
        path_b["config"] = {"workload": f"{args.etkdg_mols} drug-like pseudo-mols (20-50 heavy atoms, {args.pool} distinct) x "
                                        f"{args.confs} conformers: ETKDG embed + MMFF94 200-iter BFGS", "data": "synthetic"}

    ids_h = ids.cpu().numpy()
    n_clusters = int(cen.numel())
    assert ids_h.min() == 0 and ids_h.max() == n_clusters - 1
    sizes = np.bincount(ids_h, minlength=n_clusters)
    assert (np.diff(sizes) <= 0).all(), "cluster sizes must be non-increasing"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = unique_pairs(n) / (ms_dev * 1e-3)
    e2e = unique_pairs(n) / (ms_e2e * 1e-3)
    peak, peak_src = measured_peaks()
    kernel_ms = float(np.mean(pass_ms))
    n_edges = None
    algo_bytes = 256.0 * n + 4.0 * n  # fingerprints read once + counts written (SURVEY.md §8d "fused count" pass)
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    # integer-issue roof of the same kernel: 64 POPC per pair at 16 lanes/clk/SM (148 SMs, measured max clock)
    pairs_per_rank = unique_pairs(n) / world
    popc_rate = pairs_per_rank * words / (kernel_ms * 1e-3)

    if tensor_path:
        # dominant kernel = tcgen05 block-scaled fp4 MMA tile (kind::mxf4 over the 0/1 E2M1 expansion, unit scale
        # factors): 2 * bits ops per pair over the tiles actually visited (upper triangle). Dense fp4 issues at 4x the
        # bf16 rate on B200 (9 vs 2.25 PFLOP/s nominal), so the roof is 4 x the MEASURED bf16 throughput.
        bf16 = 1700.1
        try:
            bf16 = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            pass
        fp4 = (words * 32) % 256 == 0  # the library's own eligibility rule (tanimoto_tc.cu); else the int8 tile runs
        mult = 4.0 if fp4 else 2.0
        tiles_pairs = unique_pairs(n) / world  # + diagonal-tile overhead (< 0.1 % at 1M)
        tops = tiles_pairs * 2.0 * words * 32 / (kernel_ms * 1e-3) / 1e12
        operand_bytes = (128 + 112) * (words * 32 // 2) / (128.0 * 224.0) if fp4 else (128 + 256) * (words * 32) / (128.0 * 256.0)
        roofline = {"bound": "tensor", "achieved": tops, "peak": mult * bf16, "unit": "TFLOP/s", "frac": tops / (mult * bf16),
                    "traffic": None,
                    "kernel": ("simTensorKernel<count, fp4, cluster2> (tcgen05.mma kind::mxf4.block_scale, neighbor_pass_tc)" if fp4
                               else "simTensorKernel<count> (tcgen05.mma kind::i8, neighbor_pass_tc)"),
                    "kernel_ms": kernel_ms, "ops_per_pair": 2 * words * 32,
                    "peak_source": f"{mult:.0f} x MEASURED_PEAKS.json bf16_tflops (dense {'fp4' if fp4 else 'u8'} = {mult:.0f} x bf16 rate; of measured)",
                    "hbm_algorithmic_GBps": (n * words * 32 / (2 if fp4 else 1) + 260.0 * n) / (kernel_ms * 1e-3) / 1e9,
                    "l2_operand_bytes_per_pair": operand_bytes,
                    "l2_operand_TBps": tiles_pairs * operand_bytes / (kernel_ms * 1e-3) / 1e12,
                    "note": "exact: 0/1 products, fp32 accumulation of sums <= 4096; HBM share negligible, the operand "
                            "stream comes from L2 (TMA, column operand multicast to the CTA pair)"}
    else:
        roofline = None
    out = {
        "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "e2m1 0/1 x e2m1 0/1 -> f32 (exact integer counts), integer threshold test = the f64 predicate", "data": "synthetic",
        "config": {"workload": "1Mx1M symmetric 2048-bit Tanimoto + Butina (sim>=0.7)" if n == 1_000_000 else
                   f"{n}x{n} symmetric 2048-bit Tanimoto + Butina (sim>=0.7) [reduced size override]",
                   "n_fingerprints": n, "fp_bits": words * 32, "cutoff": CUTOFF, "pairs_counted": "unique n(n-1)/2",
                   "l2": "inputs (256 MB) larger than L2", "parallelism": f"row-group x{world}" if world > 1 else "1gpu"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(n * words * 4), "d2h_bytes_per_step": int(n * 4),
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roofline if roofline is not None else {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "simTileKernel<count> (neighbor_pass)", "kernel_ms": kernel_ms,
                     "algorithmic_bytes": algo_bytes, "peak_source": peak_src,
                     "note": "pass is integer-issue bound by construction (5e-4 B/pair); see popc_roof"},
        "popc_roof": {"achieved_popc_per_s": popc_rate, "peak_popc_per_s": 148 * 16 * 1.965e9,
                      "frac": popc_rate / (148 * 16 * 1.965e9), "unit": "32-bit POPC/s"},
        "phases_ms": phases, "n_clusters": n_clusters,
    }

    # CPU baseline (oracle port, OpenMP on the host cores) on a bounded sample + exact parity on that sample
    import oracle

    cpu_threads = oracle.set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1 to its workers
    ns, _ = cpu_sample_size(target_s=12.0)
    if args.n_centres:
        ns = min(ns, n)
    fps = synthetic.clustered_fingerprints(ns // 50, 50, seed=synthetic.SEED)
    t0 = time.perf_counter()
    ids_cpu, cen_cpu = oracle.butina_fp(fps, CUTOFF)
    dt = time.perf_counter() - t0
    g_ids, g_cen = fused_butina_device(torch.from_numpy(fps.view(np.int32)).to(dev), CUTOFF)
    parity = bool((g_ids.cpu().numpy() == ids_cpu).all() and (g_cen.cpu().numpy() == cen_cpu).all())
    out["cpu_baseline"] = {"value": unique_pairs(len(fps)) / dt, "unit": UNIT, "cores": cpu_threads,
                           "kind": "port",
                           "sample": f"{len(fps)}x{len(fps)} clustered 2048-bit fingerprints, cutoff {CUTOFF}, {dt:.1f} s"}
    out["parity_on_sample"] = "bit-exact" if parity else "MISMATCH"
    out["etkdg_mmff"] = path_b
    out["cross_similarity"] = cross
    if path_b is not None:
        flat_b, mmff_b = path_b_pool(args.pool, synthetic.SEED)
        nb = min(args.etkdg_cpu_mols, path_b["n_mols"])
        v, dt_b, okf = run_path_b_cpu(flat_b, mmff_b, nb, path_b["confs_per_mol"])
        path_b["cpu_baseline"] = {"value": v, "unit": "mols/s", "cores": cpu_threads, "kind": "port",
                                  "sample": f"{nb} mols x {path_b['confs_per_mol']} conformers, {dt_b:.1f} s, embedded {okf:.2f}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- path B (conformers)
ETKDG_PARAMS = dict(seed=20260924, boxSize=10.0, optimizerForceTol=1e-3, enforceChirality=1, useExpTorsions=1,
                    useBasicKnowledge=1, maxAttempts=0, dgIters=400, fourthIters=200, etkIters=300, maxRestarts=20)


MAX_ATTEMPTS = 20  # embedding attempts per conformer slot in the bench (the API default is 10 x atoms, src/etkdg.cpp:71-85)


def path_b_pool(pool: int, seed: int):
    """`pool` distinct pseudo drug-like molecules (20-50 heavy atoms, hydrogens added) with DG/ETK/check and MMFF tables."""
    from nvmolkit_b200 import synthetic
    from nvmolkit_b200.forcefield import FlatSystem

    flat, mols = synthetic.random_embed_molecules(pool, 20, 50, seed=seed, strict_checks=False)
    mmff = FlatSystem.from_molecules("mmff", [len(m["z"]) for m in mols], [m["terms"] for m in mols])
    return flat, mmff


def run_path_b_gpu(flat, mmff, n_mols: int, confs: int, dev, steps: int, warmup: int, world: int = 1, rank: int = 0):
    """ETKDG embed of `confs` conformers for n_mols molecules (the pool cycled), then MMFF94 200-iteration BFGS of every
    embedded conformer. Returns dict with mols/s (device-resident tables; coordinates are produced on the device)."""
    import torch
    import torch.distributed as dist

    from nvmolkit_b200 import _lib
    from nvmolkit_b200.distributed import all_gather_v, molecule_range
    from nvmolkit_b200.embedMolecules import EmbedParameters, embed_slots
    from nvmolkit_b200.forcefield import ConformerBatch
    from nvmolkit_b200.minimizer import minimize

    pool = len(flat)
    lo, hi = molecule_range(n_mols, rank, world)
    mol_ids = (np.arange(lo, hi) % pool).astype(np.int32)
    params = EmbedParameters(randomSeed=ETKDG_PARAMS["seed"])
    max_attempts = MAX_ATTEMPTS

    def step():
        raw = embed_slots(flat, params, confs, max_attempts, mol_indices=mol_ids)
        ok = raw.ok.bool()
        # MMFF on the embedded conformers (device-resident hand-over: coordinates never leave the GPU)
        batch = ConformerBatch(raw.slot_mol, raw.slot_atom_start, np.zeros((0, 3)))
        res = minimize(mmff, batch, 200, 1e-4, positions=raw.coords, active=ok.to(torch.uint8))
        if world > 1:  # the one collective of the path: all-gather of the results
            all_gather_v(res.energies)
            all_gather_v(res.positions)
        return raw, res

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        raw, res = step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = raw.ok.cpu().numpy().astype(bool)
    st = res.status.cpu().numpy()
    return {"mols_per_s": n_mols / (float(ms.item()) * 1e-3), "ms_per_step": float(ms.item()), "n_mols": n_mols,
            "confs_per_mol": confs, "conformers_embedded_frac": float(ok.mean()),
            "mean_attempts": float(raw.attempts.float().mean().item()),
            "mmff_converged_frac": float((st[ok] == 0).mean()) if ok.any() else 0.0,
            "stage_failures": raw.stage_failures.cpu().numpy().tolist(),
            "phases_ms": {"etkdg": _lib.profile_read("etkdg"), "mmff_bfgs": _lib.profile_read("bfgs")},
            "gpu_launches": int(_lib.launch_count() - l0), "atoms_per_mol_mean": float(flat.atom_counts.mean())}


def run_path_b_cpu(flat, mmff, n_mols: int, confs: int):
    """Same pipeline on the host cores with the oracle (OpenMP over conformer slots). Returns mols/s."""
    import oracle

    pool = len(flat)
    mol_ids = (np.arange(n_mols) % pool).astype(np.int32)
    slot_mol = np.repeat(mol_ids, confs)
    starts = np.concatenate([[0], np.cumsum(flat.atom_counts[slot_mol])]).astype(np.int32)
    p = dict(ETKDG_PARAMS, maxAttempts=MAX_ATTEMPTS)
    t0 = time.perf_counter()
    coords, ok, att, en = oracle.etkdg_embed_batch((flat.dg.atom_counts, flat.dg.tables), (flat.etk.atom_counts, flat.etk.tables),
                                                  flat.checks.tables, flat.checks.num_impropers, p, slot_mol, starts)
    keep = np.nonzero(ok)[0]
    if len(keep):
        k_starts = np.concatenate([[0], np.cumsum(flat.atom_counts[slot_mol[keep]])]).astype(np.int32)
        rows = np.concatenate([np.arange(starts[s], starts[s + 1]) for s in keep])
        oracle.ff_minimize("mmff", mmff.atom_counts, mmff.tables, slot_mol[keep], k_starts, coords[rows], 200, 1e-4)
    dt = time.perf_counter() - t0
    return n_mols / dt, dt, float(ok.mean())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="butina")
    ap.add_argument("--n-centres", type=int, default=0, help="override the problem size (x50 fingerprints); testing only")
    ap.add_argument("--cross-n", type=int, default=32768, help="rows of the materialised cross-similarity leg (0 = skip)")
    ap.add_argument("--etkdg-mols", type=int, default=2048, help="molecules of the ETKDG+MMFF leg (0 = skip)")
    ap.add_argument("--confs", type=int, default=10)
    ap.add_argument("--pool", type=int, default=64, help="distinct pseudo-molecules cycled to fill the batch")
    ap.add_argument("--etkdg-cpu-mols", type=int, default=16, help="molecules of the CPU-baseline sample of that leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
