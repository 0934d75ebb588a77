#!/usr/bin/env python3
"""Benchmark of the BROADCAST hot path on B200 (see BASELINE.json / SURVEY.md section 8(d)).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on host cores

Workload (config.workload): C5 = synthetic 2-D boundary layer, 8192 x 2048 cells (gh = 3, order 5),
slab-sharded in i across the N GPUs of one box (strong scaling: the global grid is fixed).
One step = halo exchange (N > 1) + the four boundary fills + one residual evaluation
(BROADCAST_npz.py:854-875, one explicit stage).  metric = FP64 residual cell-updates/s.
The Jacobian assembly of the same state is timed separately and reported in the "jacobian" object.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES_BYTES_PER_CELL = 136.0  # 12 doubles read (w5, nx2, ny2, vol, volf2) + 5 written, SURVEY.md 8(d)
JAC_BYTES_PER_CELL = 5904.0  # 29 blocks x 25 doubles written + 13 doubles read, SURVEY.md 8(d)


def kernel_source_hash():
    """sha256 over the sources of the default residual kernel: the key that ties an ncu capture to the code that ran"""
    import hashlib
    h = hashlib.sha256()
    for f in ("residual_fast.cuh", "residual_bulk.cu", "residual_tile.cu", "scheme.cuh", "grid.cuh"):
        h.update(open(os.path.join(ROOT, "broadcast_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(im, jm, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the default residual kernel, from the sidecar that
    tools/ncu_traffic.py writes next to the round's `ncu --set full` capture (profiles/residual_traffic.json).  Only reported when
    the capture was taken at this grid size AND from the kernel sources that are running now; otherwise null (a stale number is
    not a measurement)."""
    p = os.path.join(ROOT, "profiles", "residual_traffic.json")
    if world != 1 or not os.path.exists(p):
        return None, None
    try:
        d = json.load(open(p))
        if d.get("grid") == [im, jm] and d.get("source_hash") == kernel_source_hash():
            return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), d.get("capture")
    except Exception:
        pass
    return None, None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--im", type=int, default=8192)
    ap.add_argument("--jm", type=int, default=2048)
    ap.add_argument("--no-jacobian", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.index = index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------------------------
# case construction (host); slab sharding lives in broadcast_b200.sharding
# ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """pin this process (and so the first-touch placement of the pinned host buffers it allocates next) to the CPUs of the NUMA node
    the GPU hangs off: with one process per GPU the host copies of every rank otherwise come out of whatever node the launcher
    started on and share its memory controllers.  Best effort: returns the node or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def build_global_case(im, jm, f_geom):
    from broadcast_b200 import cases
    return cases.make_bl_case(im, jm, f_geom=f_geom, name=f"bl2d_{im}x{jm}")


# ----------------------------------------------------------------------------------------------
# CPU side: the reference's own Fortran (machine-translated to C, oracle/_ref) on host cores
# ----------------------------------------------------------------------------------------------
def cpu_sample_worker(args):
    im, jm, steps, warm = args
    from oracle import refmods
    from broadcast_b200 import cases
    R = refmods.make(fast=True)
    c = cases.make_bl_case(im, jm, f_geom=R["f_geom"])
    w = c.w.copy(order="F")
    res = c.zeros_state()
    for _ in range(warm):
        cases.apply_bcs(c, w, R["f_bnd"])
        R["f_sch"].flux_num_dnc5_2d(res, w, *c.scheme_args())
    t0 = time.perf_counter()
    for _ in range(steps):
        cases.apply_bcs(c, w, R["f_bnd"])
        R["f_sch"].flux_num_dnc5_2d(res, w, *c.scheme_args())
    dt = time.perf_counter() - t0
    return im * jm * steps / dt, dt / steps


def cpu_jacobian_colour_seconds(sample=(1024, 512), colours=2):
    """reference colour pass (seed + 4 linearised boundary fills + tangent; the COO scatter of the reference layout does not
    fit at this size) on one host core: seconds per colour per cell"""
    from oracle import refmods
    from broadcast_b200 import cases
    R = refmods.make(fast=True)
    im, jm = sample
    c = cases.make_bl_case(im, jm, f_geom=R["f_geom"])
    w = c.w.copy(order="F")
    cases.apply_bcs(c, w, R["f_bnd"])
    wd = c.zeros_state()
    res, resd = c.zeros_state(), c.zeros_state()
    t0 = time.perf_counter()
    for n in range(colours):
        wd *= 0.0
        R["f_misc"].testvector(wd, n % 5, n % 7, (3 * n) % 7, c.gh, im, jm)
        cases.apply_bcs_lin(c, w, wd, R["f_bnd"], R["f_lin"])
        R["f_lin"].flux_num_dnc5_2d_d(res, resd, w, wd, *c.scheme_args())
    return (time.perf_counter() - t0) / colours / (im * jm)


def cpu_baseline(sample=(1024, 512), steps=8, warm=1):
    v, ms = cpu_sample_worker((sample[0], sample[1], steps, warm))
    return {"value": v, "unit": "cell-updates/s", "cores": 1, "kind": "reference",
            "sample": f"{steps} steps (4 boundary fills + residual) of the C5 recipe at {sample[0]}x{sample[1]} cells, reference Fortran "
                      f"machine-translated to C (oracle/_ref, gcc -O3), 1 thread as the reference is serial"}


def run_reference(a):
    """The reference's own CPU path (oracle/_ref: its Fortran machine-translated to C, gcc -O3 -march=x86-64-v3) on the SAME
    config as the CUDA arm: ONE replica of the full a.im x a.jm grid (the reference is serial code: one block = one core),
    min(steps, 2) timed steps of [4 boundary fills + residual].  The all-core figure (one bounded replica per host core, what a
    user with many independent cases could extract from the box) is kept next to it as cpu_baseline.all_cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    steps = max(1, min(a.steps, 2))
    warm = 1 if a.warmup > 0 else 0
    t0 = time.perf_counter()
    value, sec = cpu_sample_worker((a.im, a.jm, steps, warm))
    wall_full = time.perf_counter() - t0
    sample = (1024, 512)
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        outs = pool.map(cpu_sample_worker, [(sample[0], sample[1], 4, 1)] * cores)
    wall_all = time.perf_counter() - t0
    allc = float(sum(o[0] for o in outs))
    line = {
        "impl": "reference", "metric": "fp64_residual_cell_updates_per_s", "value": float(value), "unit": "cell-updates/s", "n_gpus": a.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5 synthetic 2-D boundary layer {a.im}x{a.jm}, order 5 (gh=3), i-slabs over {a.gpus} GPU(s)",
                   "step": "halo exchange (N>1) + 4 boundary fills + 1 residual (flux_num_dnc5_2d)", "l2": "n/a (CPU)",
                   "note": "full grid, one replica: the reference is serial Fortran (no OpenMP / MPI inside a block)"},
        "cpu_baseline": {"value": float(value), "unit": "cell-updates/s", "cores": 1, "kind": "reference",
                         "sample": f"{steps} steps (4 boundary fills + residual) on the full {a.im}x{a.jm} grid, reference Fortran machine-translated "
                                   f"to C (oracle/_ref, gcc -O3 -march=x86-64-v3), 1 thread as the reference is serial; wall {wall_full:.1f} s",
                         "all_cores": {"value": allc, "cores": cores,
                                       "sample": f"{cores} independent replicas, one per host core, each 4 steps at {sample[0]}x{sample[1]} cells "
                                                 f"(cache-sized: favours the CPU); wall {wall_all:.1f} s"}},
        "e2e": {"value": float(value), "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    # rank 0 prints exactly ONE line on stdout: anything a library writes to fd 1 on the way (NCCL prints its version there when
    # NCCL_DEBUG is set) goes to stderr instead, and the JSON line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (line + "\n").encode())
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import ctypes
    import broadcast_b200 as bb
    from broadcast_b200 import _lib
    from broadcast_b200.resident import Block, _p

    from broadcast_b200 import sharding
    from broadcast_b200.resident import jacobian_hybrid

    peak, peak_kind = measured_peaks()
    gcase = build_global_case(a.im, a.jm, bb.f_geom)
    case, slab = sharding.slab_of(gcase, rank, world)
    blk = Block(case, dev, slab=slab if world > 1 else None)
    gcase_holder = [gcase]
    del gcase
    # everything below runs on a stream of its own: a CUDA graph cannot be captured on (or replayed into) the legacy default
    # stream without extra cross-stream waits
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    # halo exchange: peer stores over NVLink (csrc/halo.cu, two launches, no NCCL / torch op on the data path); NCCL send/recv
    # (sharding.HaloExchange) only if the mailboxes cannot be mapped (BROADCAST_B200_HALO=nccl forces it)
    halo_kind = "none"
    halo = lambda w: None
    if world > 1:
        halo_kind = "peer-store kernels over NVLink (csrc/halo.cu)"
        try:
            if os.environ.get("BROADCAST_B200_HALO", "peer") != "peer":
                raise RuntimeError("forced")
            halo = sharding.PeerHalo(case.gh, rank, world, blk.w)
        except Exception as e:   # noqa: BLE001
            halo = sharding.HaloExchange(case.gh, rank, world)
            halo_kind = f"NCCL send/recv (peer mailboxes unavailable: {e})"
    gcase_e2e = gcase_holder[0] if not a.no_e2e else None
    del gcase_holder[:]
    cells_global = a.im * a.jm
    cells_local = case.im * case.jm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # BROADCAST_B200_OVERLAP=1: halo exchange + boundary fills on a side stream while the inner tiles of the residual run
    # (Block.step_overlapped).  Measured on B200 (profiles/r1_g_summary.md): no gain -- N = 1: 2.40 vs 2.35 ms, N = 8: 0.378 vs 0.375 ms
    # per step (the one-wave ring launch and the stream joins cost what the overlap hides) -- so the plain sequence is the default.
    overlap = os.environ.get("BROADCAST_B200_OVERLAP", "0") == "1"
    # the step [exchange, fills, residual] as ONE captured CUDA graph (sharding.StepGraph): at N = 8 the kernel takes 0.3 ms and
    # the Python-issued launches were a fixed ~70 us per step; BROADCAST_B200_STEP_GRAPH=0 issues the calls one by one
    # (one GPU: the 2.4 ms kernel hides every launch; the calls stay separate so that CUDA events bracket the kernel inside the
    # timed region)
    use_graph = os.environ.get("BROADCAST_B200_STEP_GRAPH", "1" if world > 1 else "0") == "1" and not overlap
    sgraph = sharding.StepGraph(blk, halo if world > 1 else None) if use_graph else None

    def step():
        if sgraph is not None:
            sgraph()
        elif overlap:
            blk.step_overlapped(halo if world > 1 else None)
        else:
            halo(blk.w)
            blk.apply_bcs()
            blk.residual()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    ev0.record()
    for s in range(a.steps):
        if sgraph is not None:
            step()
        elif overlap:
            kev[s][0].record()
            step()
            kev[s][1].record()
        else:
            halo(blk.w)
            blk.apply_bcs()
            kev[s][0].record()
            blk.residual()
            kev[s][1].record()
    ev1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms_total = ev0.elapsed_time(ev1)
    kernel_timing = "CUDA events around the kernel launch of every timed step"
    if sgraph is not None:
        # the graph has no place for events around its kernel node: the same kernel on the same state, K launches right after
        for s in range(a.steps):
            kev[s][0].record()
            blk.residual()
            kev[s][1].record()
        barrier()
        kernel_timing = "K launches of the kernel alone right after the K timed graph replays (same state)"
    k_ms = float(np.mean([x.elapsed_time(y) for x, y in kev]))
    t = torch.tensor([ms_total, k_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, k_ms = float(t[0]), float(t[1])
    ms_step = ms_total / a.steps
    value = cells_global / (ms_step * 1e-3)

    # checksum of the step's result over the OWNED cells of all ranks: wrap-around int64 sum of the bit patterns of the residual
    # (order independent, so identical for every N iff the sharded residual equals the single-GPU one bit for bit) + L2 norms
    gh_ = case.gh
    own = blk.res[:, gh_:gh_ + case.jm, gh_:gh_ + case.im].contiguous()
    bits = own.view(torch.int64).sum().reshape(1)
    sq = (own * own).sum(dim=(1, 2))
    if world > 1:
        dist.all_reduce(bits, op=dist.ReduceOp.SUM)
        dist.all_reduce(sq, op=dist.ReduceOp.SUM)
    checksum = {"res_bits_sum_i64": int(bits.item()), "res_l2": [float(x) for x in sq.sqrt().cpu()],
                "note": "wrap-around int64 sum of the residual's bit patterns over the owned cells of all ranks (order independent): "
                        "equal across N = 1/2/4/8 iff the slab-sharded residual is bit-identical to the single-block one"}
    del own, bits, sq

    # roofline of the dominant kernel (fused residual tile kernel), per launch, per GPU
    achieved = RES_BYTES_PER_CELL * cells_local / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(a.im, a.jm, world)
    roofline = {"bound": "hbm", "kernel": "k_residual_fast_bulk (32x9 tile, 320 threads, every input of the tile by TMA)" + ("; inner tiles + ring of tiles = 2 launches per step overlapping the halo exchange and boundary fills, timed first launch to end of last" if overlap else ""), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_kind": peak_kind, "kernel_timing": kernel_timing, "traffic": traffic, "traffic_capture": traffic_src, "kernel_ms": k_ms,
                "algorithmic_bytes_per_cell": RES_BYTES_PER_CELL,
                "fp64_pipe_note": "970 FP64 instructions per cell (0.87 ms at 100 % of the FP64 pipe): on-chip bound, see DESIGN.md section 4 and profiles/r2_b_summary.md"}

    # Jacobian assembly of the same state (BASELINE.json metric, second half): regular rows by face linearisation into the
    # fixed 29-block pattern + the four boundary strips by the reference colour loop; algorithmic bytes 5904 B per cell
    jac = None
    if not a.no_jacobian:
        nblk = 29 * 25 * cells_local * 8
        free, _tot = torch.cuda.mem_get_info(dev)
        if nblk + 40 * blk.w.numel() * 8 < 0.9 * free:
            blocks = torch.empty((29, 5, 5, case.jm, case.im), dtype=torch.float64, device=dev)
            cd = torch.zeros((case.jm, case.im), dtype=torch.float64, device=dev)
            halo(blk.w)
            blk.apply_bcs()
            jacobian_hybrid(blk, coefdiag=cd, blocks=blocks)   # warm-up (allocations, first launches)
            barrier()
            nrep = 3
            j0, j1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # (the halo columns and ghosts of the state are current: the assembly itself exchanges nothing, and an exchange inside
            #  the timed loop would add the skew between the ranks' host loops to a device time)
            halo(blk.w)
            blk.apply_bcs()
            barrier()
            j0.record()
            for _ in range(nrep):
                i0.record()
                blk.call("bcd_jacobian_interior", _p(blocks), _p(blk.w), _p(blk.nx), _p(blk.ny), _p(blk.vol), _p(blk.volf), blk.gh, *blk._phys,
                         case.im, case.jm, _p(cd), ctypes.c_void_p(None), blk._stream())
                i1.record()
                H = jacobian_hybrid(blk, coefdiag=cd, blocks=blocks)
            j1.record()
            barrier()
            # the interior pass is timed once more inside the loop (i0..i1) to report its share; subtract it from the total
            int_ms = i0.elapsed_time(i1)
            tot_ms = j0.elapsed_time(j1) / nrep - int_ms
            tt = torch.tensor([tot_ms, int_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tot_ms, int_ms = float(tt[0]), float(tt[1])
            ach = JAC_BYTES_PER_CELL * cells_local / (tot_ms * 1e-3) / 1e9
            jac = {"assembly_s": tot_ms * 1e-3, "cells": cells_global, "cells_per_s": cells_global / (tot_ms * 1e-3),
                   "interior_blocks_ms": int_ms, "strips_and_fill_ms": tot_ms - int_ms,
                   "layout": "29 fixed 5x5 blocks per cell (values[slot][25][cell]) + COO strips in the reference's slot order",
                   "roofline": {"bound": "hbm", "kernel": "k_face_packages + k_jac_assemble_rt (+ strip colour loops)", "achieved": ach,
                                "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_cell": JAC_BYTES_PER_CELL,
                                "interior_only_frac": JAC_BYTES_PER_CELL * cells_local / (int_ms * 1e-3) / 1e9 / peak}}
            # CSR row block of this rank (zero filter, reference numbering, columns ascending: csrc/csr.cu) where it fits next
            # to the block values (C5 on one GPU: 97 GB of blocks + ~75 GB of CSR do not)
            free, _tot = torch.cuda.mem_get_info(dev)
            if 110.0 * 5 * cells_local * 12 < 0.8 * free:
                # (the index / value arrays of the first conversion are overwritten by the timed one, as successive Newton
                #  iterations do: allocating tens of GB costs more than the conversion itself)
                H.to_csr(slack=True)
                store = H.csr_storage
                barrier()
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                ip, _idx, _dat = H.to_csr(out=store)
                c1.record()
                barrier()
                cm = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
                nz = torch.tensor([float(ip[-1].item())], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(cm, op=dist.ReduceOp.MAX)
                    dist.all_reduce(nz, op=dist.ReduceOp.SUM)
                jac["csr_ms"] = float(cm[0])
                jac["csr_nnz"] = int(nz[0])
                jac["assembly_plus_csr_s"] = jac["assembly_s"] + float(cm[0]) * 1e-3
                if world > 1:
                    # the per-GPU row blocks gathered on the host for the reference's PETSc path (north_star): device-to-host copy of
                    # this rank's (indptr, indices, data) through a reused pinned staging buffer of 1 GiB, max over ranks
                    stage = torch.empty(1 << 27, dtype=torch.float64).pin_memory()
                    nbytes = 0
                    barrier()
                    t0 = time.perf_counter()
                    for t_ in (ip, _idx, _dat):
                        flat = t_.view(torch.uint8).view(-1)
                        sb_ = stage.view(torch.uint8)
                        for a0 in range(0, flat.numel(), sb_.numel()):
                            n_ = min(sb_.numel(), flat.numel() - a0)
                            sb_[:n_].copy_(flat[a0:a0 + n_], non_blocking=True)
                            torch.cuda.current_stream().synchronize()
                        nbytes += flat.numel()
                    gs = torch.tensor([time.perf_counter() - t0, float(nbytes)], dtype=torch.float64, device=dev)
                    gmax = gs.clone()
                    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
                    dist.all_reduce(gs, op=dist.ReduceOp.SUM)
                    jac["gather_s"] = float(gmax[0])
                    jac["gather_bytes"] = int(gs[1])
                    del stage
                del ip, _idx, _dat, store
                del blocks, H
                torch.cuda.empty_cache()
            else:
                # C5 on one GPU: 97 GB of block values + 73 GB of CSR do not fit together.  The contract output (CSR / vol, what
                # misc/PETSc_func.py:71-95 consumes) comes from the banded assembly: i-bands through ONE reused band buffer of
                # block values, the zero filter and the CSR conversion per band (resident.BandedAssembly); timed as a whole.
                del blocks, H
                torch.cuda.empty_cache()
                jac["csr_ms"] = None
                if world == 1:
                    from broadcast_b200.resident import BandedAssembly
                    nband = int(os.environ.get("BROADCAST_B200_JAC_BANDS", "8"))
                    ba = BandedAssembly(case, nband, dev)
                    parts = ba.assemble_csr(blk.w)           # warm-up: allocations, graph captures
                    nnz = int(sum(int(p_[0][-1].item()) for p_ in parts))
                    del parts                                # the CSR arrays stay with `ba`: the next assembly overwrites them
                    barrier()
                    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    b0.record()
                    parts = ba.assemble_csr(blk.w)
                    b1.record()
                    barrier()
                    jac["assembly_plus_csr_s"] = b0.elapsed_time(b1) * 1e-3
                    jac["csr_nnz"] = nnz
                    jac["csr_bands"] = nband
                    jac["csr_note"] = (f"banded: {nband} i-bands through one reused band buffer of block values ({ba._buf.numel() * 8 / 2**30:.1f} GiB); "
                                       "Jacobian -> zero filter -> CSR / vol, row blocks in row order, index / value arrays of the previous assembly reused")
                    jac["roofline"]["assembly_plus_csr_frac"] = JAC_BYTES_PER_CELL * cells_local / jac["assembly_plus_csr_s"] / 1e9 / peak
                    del parts, ba
                    torch.cuda.empty_cache()
        else:
            jac = {"skipped": "block values do not fit next to the state on this GPU at this size"}
    clocks = sampler.stop() if rank == 0 else {}

    # end to end through the plugin-level call with pinned HOST buffers (state in, residual out, every step)
    e2e = None
    if not a.no_e2e:
        numa = bind_to_gpu_numa_node(local)
        wp = torch.empty(blk.w.shape, dtype=torch.float64).pin_memory()
        rp = torch.empty(blk.w.shape, dtype=torch.float64).pin_memory()
        wp.copy_(blk.w.cpu())
        nst = max(3, min(a.steps, 10))
        # the host step is pipelined over i-slabs of the rank's part of the block (H2D of slab k+1 / kernels of slab k / D2H of
        # slab k-1 overlap on three streams).  Every slab's gh halo columns come straight from the HOST array (which holds the
        # rank's columns plus its halo columns, as a rank of the reference's host would hold them), so no device exchange is needed.
        from broadcast_b200.resident import StreamedBlock
        nslab = int(os.environ.get("BROADCAST_B200_E2E_SLABS", "16" if world == 1 else "4"))
        # one GPU: slab widths doubling from both ends towards the middle, so that the first device-to-host copy starts after a
        # short upload and the step ends with a short download (BROADCAST_B200_E2E_TAPER=0: even slabs)
        bounds = None
        rows_mode = world == 1 and os.environ.get("BROADCAST_B200_E2E_ROWS", "1") == "1"
        if rows_mode:
            # one GPU: pipeline over ROW windows, every host-link copy one contiguous run per plane (resident.RowStreamedBlock;
            # the pitched copies of the i-slab pipeline reach 29.6 GB/s each way with both directions busy, contiguous ones 43)
            from broadcast_b200.resident import RowStreamedBlock
            sb = RowStreamedBlock(gcase_e2e, nwin=nslab, device=dev)
            api = ("broadcast_b200.resident.RowStreamedBlock.step_from_host (pinned host w in, residual out, " + str(nslab) +
                   " pipelined row windows with contiguous host-link copies; mesh metrics resident)")
        else:
            if world == 1 and os.environ.get("BROADCAST_B200_E2E_TAPER", "1") == "1":
                bounds = StreamedBlock.tapered_bounds(gcase_e2e.im, nslab)
            sb = StreamedBlock(gcase_e2e, nslab=nslab * world, device=dev, first=rank * nslab, count=nslab, bounds=bounds)
            api = ("broadcast_b200.resident.StreamedBlock.step_from_host per rank (pinned host w in, residual out, " + str(nslab) +
                   " pipelined i-slabs per GPU" + (", widths tapered towards both ends" if bounds else "") + "; mesh metrics resident)")
        run = lambda: sb.step_from_host(wp, rp)
        h2d, d2h = sb.bytes_per_step()
        for _ in range(2):
            run()
        barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            run()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / nst], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": cells_global / float(dt[0]), "unit": "cell-updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": float(dt[0]) * 1e3, "api": api, "host_numa_node": numa}
        if world == 1:
            # what came back through the host buffers is the resident step's residual, bit for bit (same checksum as above)
            gh_ = case.gh
            hb = rp[:, gh_:gh_ + case.jm, gh_:gh_ + case.im].contiguous().view(torch.int64).sum()
            e2e["result_bits_sum_i64"] = int(hb.item())
            e2e["result_matches_resident"] = bool(int(hb.item()) == checksum["res_bits_sum_i64"])

    if rank == 0:
        line = {
            "metric": "fp64_residual_cell_updates_per_s", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5 synthetic 2-D boundary layer {a.im}x{a.jm}, order 5 (gh=3), i-slabs over {world} GPU(s)",
                       "step": "halo exchange (N>1) + 4 boundary fills + 1 residual (flux_num_dnc5_2d)" + (", fills and exchange overlapped with the inner tiles" if overlap else ""),
                       "halo": halo_kind, "step_issue": (("one CUDA-graph replay per step" + (", exchange + fills forked beside the inner residual tiles" if getattr(sgraph, "overlap", False) else ""))
                                      if sgraph is not None else "separate launches"),
                       "l2": f"inputs larger than L2 ({blk.w.numel() * 8 / 2**20:.0f} MiB state per GPU)"},
            "roofline": roofline, "jacobian": jac, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "checksum": checksum,
        }
        if not a.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
            if jac is not None and "assembly_s" in jac:
                spc = cpu_jacobian_colour_seconds()
                jac["cpu_baseline"] = {"value": 245 * spc * cells_global, "unit": "s", "cores": 1, "kind": "reference",
                                       "sample": "2 colour passes (seed + 4 linearised boundary fills + tangent, no COO scatter) of the C5 "
                                                 "recipe at 1024x512 on oracle/_ref, extrapolated to 245 colours x 8192x2048 cells: the "
                                                 "reference's COO layout (329 GB, int32 slots) cannot hold C5"}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
