#!/usr/bin/env python
"""Benchmark of the hot path: yk_solution::run_solution for iso3dfd (16th order, fp32).

Contract (see task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line
on rank 0.  Metric = GPoints/s (= domain points x steps / seconds inside run_solution, the reference's
"throughput (num-points/sec)", /root/reference/src/kernel/lib/soln_apis.cpp:455-461).

  value       whole-job GPoints/s with all inputs resident in HBM, device time (CUDA events on the
              launching stream), max over ranks.
  e2e         same metric through the public C-ABI with HOST buffers: set_elements_in_slice (H2D from
              pinned memory) of p(t), p(t-1), v  ->  run_solution(K steps)  ->  get_elements_in_slice
              (D2H) of the final p.  A time-stepping job moves its state once, not every step, so the
              per-step byte counts are total/K.
  roofline    algorithmic 16 B/point-step (read p(t), p(t-1), v; write p(t+1)) x points per launch
              / mean kernel launch time, against the measured copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline the reference's own optimized CPU path (oracle/_ref, built from the unmodified
              sources) timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU implementation instead (rank 0 only).
N > 1: weak scaling, 1024^3 points per GPU, domain split along x (the outermost storage dim).
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_POINT = 16  # 3 arrays read + 1 written, fp32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=1024, help="points per dim per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference CPU arm / cpu_baseline
# --------------------------------------------------------------------------------------------------
def host_cpu_info():
    model, flags = "unknown", ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and model == "unknown":
                model = line.split(":", 1)[1].strip()
            if line.startswith("flags") and not flags:
                flags = line
    except OSError:
        pass
    return model, flags, os.cpu_count() or 1


def run_reference_cpu(size, steps, trials=1):
    """Time the unmodified reference's optimized path (oracle/_ref/yask/bin/yask_kernel.iso3dfd.<arch>.exe,
    the reference's own harness src/kernel/yask_main.cpp) on all host cores.  Falls back to the C oracle
    port (oracle/yask_oracle.c, OpenMP) if the prebuilt reference cannot run on this CPU."""
    model, flags, ncores = host_cpu_info()
    bind = os.path.join(ROOT, "oracle", "_ref", "yask", "bin")
    libd = os.path.join(ROOT, "oracle", "_ref", "yask", "lib")
    archs = [a for a in (("avx512" if "avx512f" in flags else None), "avx2") if a]
    env = dict(os.environ, LD_LIBRARY_PATH=libd + ":" + os.environ.get("LD_LIBRARY_PATH", ""), OMP_NUM_THREADS=str(ncores),
               OMP_PLACES="cores")
    for arch in archs:
        exe = os.path.join(bind, f"yask_kernel.iso3dfd.{arch}.exe")
        if not os.path.exists(exe):
            continue
        cmd = [exe, "-g", str(size), "-no-pre_auto_tune", "-no-auto_tune", "-no-warmup", "-num_trials", str(trials), "-trial_steps",
               str(steps), "-sleep", "0", "-no-print_suffixes"]
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        except Exception:
            continue
        m = re.findall(r"best-throughput \(num-points/sec\):\s*([0-9.eE+]+)", r.stdout)
        th = re.findall(r"Num OpenMP threads used:\s*(\d+)", r.stdout)
        if r.returncode == 0 and m:
            return dict(value=float(m[-1]) / 1e9, unit="GPoints/s", cores=int(th[-1]) if th else ncores, kind="reference",
                        sample=f"iso3dfd r=8 fp32 {size}^3 x {steps} steps x {trials} trial(s), reference yask_kernel.iso3dfd.{arch}.exe "
                               f"(g++ -O3, OpenMP, BKC blocks 96x28x96), CPU: {model}")
    # fallback: oracle port
    import numpy as np
    from oracle import oracle as O
    from yask_b200.synth import hash_field, var_salt
    n = min(size, 256)
    p0 = hash_field(1, var_salt("p", 0), (-8, -8, -8), (n + 16,) * 3, -1, 1)
    p1 = hash_field(1, var_salt("p", 1), (-8, -8, -8), (n + 16,) * 3, -1, 1)
    v = hash_field(1, var_salt("v", 0), (0, 0, 0), (n,) * 3, 0.05, 0.3)
    t0 = time.time()
    O.iso3dfd_run(p0, p1, v, 8, steps, 2)
    dt = time.time() - t0
    return dict(value=n ** 3 * steps / dt / 1e9, unit="GPoints/s", cores=O.lib().yo_num_threads(), kind="port",
                sample=f"iso3dfd r=8 fp32 {n}^3 x {steps} steps, oracle/yask_oracle.c (OpenMP), CPU: {model}")


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" = one reference time-step on a bounded sample of the workload
    size = args.size if args.size <= 1024 else 1024
    total_steps = max(1, args.steps)
    if args.warmup:
        run_reference_cpu(min(size, 256), max(1, min(args.warmup, 3)))
    t0 = time.time()
    cb = run_reference_cpu(size, total_steps)
    wall = time.time() - t0
    pts = size ** 3
    line = {"metric": f"GPoints/s, iso3dfd-16 fp32 {size}^3 per GPU", "value": cb["value"], "unit": "GPoints/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": pts / (cb["value"] * 1e9) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"iso3dfd r=8 fp32 {size}^3 on the host CPU (reference OpenMP/AVX path)", "wall_s": round(wall, 1)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "GPoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def main_b200(args):
    import numpy as np
    import torch
    from yask_b200 import capi
    from yask_b200.synth import var_salt

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("YB_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        import torch.distributed as dist_
        dist = dist_
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    N = args.size
    K, W = args.steps, max(args.warmup, 3)

    s = capi.Solution("iso3dfd")
    s.set_rank_domain_size_vec([N, N, N])
    if world > 1:
        s.set_num_ranks_vec([world, 1, 1])
        s.set_rank_index_vec([rank, 0, 0])
    for kv in args.opt:
        k, v = kv.split("=", 1)
        s.set_option(k, v)
    s.prepare_solution(local)
    if world > 1:
        from yask_b200 import multi
        multi.connect(s, dist, rank, world)
    p, v = s.get_var("p"), s.get_var("v")
    seed = 2024
    for t in (0, 1):
        p.fill_hash(t, seed, var_salt("p", t), -1.0, 1.0)
    v.fill_hash(0, seed, var_salt("v", 0), 0.05, 0.3)
    s.sync()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then the timed region: exactly K steps --------------------------------------------
    s.run_solution(0, W - 1)
    barrier()
    s.clear_stats()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.25)
    barrier()
    t0 = time.time()
    s.run_solution(W, W + K - 1)
    s.sync()
    barrier()
    t1 = time.time()
    st = s.get_stats()
    clk = clocks.stop(t0, t1)
    dev_s = st.elapsed_secs                         # CUDA events on the launching stream
    wall_s = t1 - t0
    launches = st.kernel_launches
    if dist:
        tt = torch.tensor([dev_s, wall_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, wall_s = tt.tolist()
    pts_per_gpu = N ** 3
    value = pts_per_gpu * world * K / dev_s / 1e9
    checksum = p.checksum(p.get_last_valid_step_index())

    # ---- roofline of the dominant kernel (the point-update kernel: one launch per step per GPU) ------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kern_ms = dev_s / K * 1e3                       # one dominant launch per step; multi-GPU steps add thin boundary launches
    achieved = BYTES_PER_POINT * pts_per_gpu / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "iso3dfd_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "kernel": "iso3dfd_tma2_kernel", "kernel_ms": round(kern_ms, 4),
                "algorithmic_bytes_per_launch": BYTES_PER_POINT * pts_per_gpu}

    # ---- e2e through the C ABI with host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
      try:
            f0, l0 = p.halo_box(0)
            shape = [b - a + 1 for a, b in zip(f0[1:], l0[1:])]
            tl = p.get_last_valid_step_index()
            # host copies of the inputs in pinned memory (set-up, untimed): read back the synthetic fields
            hp = [torch.empty(shape, dtype=torch.float32, pin_memory=True) for _ in range(2)]
            hv = torch.empty([N, N, N], dtype=torch.float32, pin_memory=True)
            hout = torch.empty([N, N, N], dtype=torch.float32, pin_memory=True)
            import ctypes as C
            L = capi.lib()

            def get_into(var, tens, first, last):
                n = C.c_int64(0)
                capi._chk(L.yb_var_get_slice(s._h, var.index, C.c_void_p(tens.data_ptr()), capi._arr(first), capi._arr(last), C.byref(n)))

            def set_from(var, tens, first, last):
                n = C.c_int64(0)
                capi._chk(L.yb_var_set_slice(s._h, var.index, C.c_void_p(tens.data_ptr()), capi._arr(first), capi._arr(last), C.byref(n)))

            for i, t in enumerate((tl - 1, tl)):
                f, l = p.halo_box(t)
                get_into(p, hp[i], f, l)
            fv, lv = v.halo_box(0)
            get_into(v, hv, fv, lv)
            barrier()
            te0 = time.time()
            for i, t in enumerate((tl - 1, tl)):           # H2D: both step slots of p, and v
                f, l = p.halo_box(t)
                set_from(p, hp[i], f, l)
            set_from(v, hv, fv, lv)
            s.run_solution(tl, tl + K - 1)                   # K steps
            tl2 = p.get_last_valid_step_index()
            get_into(p, hout, *p.domain_box(tl2))            # D2H of the result (syncs)
            s.sync()
            barrier()
            te1 = time.time()
            e2e_s = te1 - te0
            if dist:
                tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e2e_s = tt.item()
            h2d = (2 * hp[0].numel() + hv.numel()) * 4
            d2h = hout.numel() * 4
            e2e = {"value": round(pts_per_gpu * world * K / e2e_s / 1e9, 2), "unit": "GPoints/s", "h2d_bytes_per_step": h2d // K,
                   "d2h_bytes_per_step": d2h // K, "seconds": round(e2e_s, 4), "h2d_bytes_total_per_gpu": h2d, "d2h_bytes_total_per_gpu": d2h,
                   "note": "set_elements_in_slice(p t-1,t; v) from pinned host + run_solution(K) + get_elements_in_slice(p)"}
            launches_e2e = s.get_stats().kernel_launches - launches
      except Exception as ex:   # keep the bench line alive (e.g. pinned-memory limits on a shared host)
        e2e = {"value": None, "unit": "GPoints/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": repr(ex)[:200]}
    s.close()

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = run_reference_cpu(1024 if N >= 1024 else N, 8)
        except Exception as e:  # keep the bench line alive
            cpu = {"value": None, "unit": "GPoints/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        line = {"metric": f"GPoints/s, iso3dfd-16 fp32 {N}^3 per GPU", "value": round(value, 2), "unit": "GPoints/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": round(dev_s / K * 1e3, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"iso3dfd radius 8 (16th order) fp32, {N}^3 points per GPU, rank grid {world}x1x1",
                           "fp_mode": "ref_gcc (bit-exact vs reference default build)", "l2": "inputs (13.5 GB/GPU) larger than L2; no flush needed",
                           "global_points": pts_per_gpu * world, "wall_ms_per_step": round(wall_s / K * 1e3, 4), "checksum": str(checksum)},
                "hbm_gbs_algorithmic": round(value * BYTES_PER_POINT, 1), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
