#!/usr/bin/env python3
"""bench.py -- Mpixels/s encode+decode of the HTJ2K hot path.

A step = one pass of the hot path over one batch of frames: each frame is encoded to a codestream,
then that codestream is decoded back (the metric BASELINE.json names is "Mpixels/s encode+decode").
Workload at every N: synthetic 8192x8192 3-component 12-bit frames, reversible 5/3 + RCT, 5 levels,
64x64 blocks (the headline configuration), OJB_BENCH_WORKERS (default 8) frames per GPU per step, each on its
own codec object / CUDA stream so that the host phases (packet headers) and copies of one frame
overlap the kernels of another (weak scaling; frames are independent, so there is no data-path
collective -- only the final gather of the codestream sizes to rank 0).

  value : frame already resident in HBM when the timed region starts (encoder's device image
          buffer), codestream left on the device, decoded image left on the device
  e2e   : the reference-facing C-ABI frame calls with HOST buffers (pinned): H2D of the planes,
          D2H of the codestream, H2D of the codestream, D2H of the decoded planes, all timed
  --impl reference : the unmodified reference (oracle/_ref, compiled from /root/reference by
          oracle/Makefile) on the host cores, one process per core, in-memory files
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PROFILE_JSON = "r01k_kernels.json"      # the committed ncu --set full summary the roofline's `traffic` comes from
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W = H = 8192
NC, BD, LEVELS = 3, 12, 5
CPU_TILE = 4096            # the CPU arms process a bounded sample: one 4096x4096 quarter frame


def workload_params(w=W, h=H):
    import openjph_b200 as ob
    return ob.make_params(w, h, NC, BD, num_decomps=LEVELS, reversible=True, color_transform=True)


def make_frame(w, h, seed):
    import images
    return [p.astype("uint16") for p in images.synth_frame(w, h, NC, BD, seed)]


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.maxc = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0])); self.maxc = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.maxc, "reasons": sorted(self.reasons)}


_CPU_FRAME = None


def cpu_worker_init(seed):
    """per-process set-up (outside the timed region): load the reference build, synthesise the frame"""
    global _CPU_FRAME
    import numpy as np
    import refharness  # noqa: F401  (dlopen of oracle/_ref)
    _CPU_FRAME = [f.astype(np.int32) for f in make_frame(CPU_TILE, CPU_TILE, seed)]


def cpu_worker(args):
    """one reference encode+decode of a CPU_TILE^2 frame; returns seconds (enc, dec) and the size"""
    seed, check = args
    import numpy as np
    import refharness as R
    if _CPU_FRAME is None:
        cpu_worker_init(seed)
    p = workload_params(CPU_TILE, CPU_TILE)
    t0 = time.perf_counter(); cs = R.encode(p, _CPU_FRAME); t1 = time.perf_counter()
    out, _ = R.decode(cs); t2 = time.perf_counter()
    if check:
        assert all(np.array_equal(a, b) for a, b in zip(out, _CPU_FRAME))
    return t1 - t0, t2 - t1, len(cs)


def run_cpu_reference(procs, steps, warmup):
    """all host threads: `procs` persistent processes, each one encodes+decodes one quarter frame
    per step; frames are synthesised once per process before the timed region"""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=cpu_worker_init, initargs=(1234,)) as pool:
        for _ in range(max(1, warmup)):
            pool.map(cpu_worker, [(1234, True)] * procs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(cpu_worker, [(1234, False)] * procs, chunksize=1)
        dt = time.perf_counter() - t0
    pix = procs * steps * CPU_TILE * CPU_TILE
    return pix / dt / 1e6, dt / steps * 1e3


def bind_to_gpu_numa_node(torch, local):
    """run this rank's host threads (and first-touch its pinned buffers) on the CPUs next to its GPU:
    pinned memory on the far socket halves the PCIe rate.  Returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d CPUs local to GPU %s" % (len(cpus), bus)
    except Exception as e:          # no NVML / no topology information: run unbound
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def main():
    # the contract is ONE JSON line on stdout: keep the real stdout for it and send everything libraries
    # print at C level (e.g. NCCL's version banner) to stderr
    out = os.fdopen(os.dup(1), "w")
    sys.stdout.flush()
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cfg = {"workload": "8192x8192x3 12-bit, reversible 5/3 + RCT, 5 levels, 64x64 code-blocks, RPCL; a step = a batch of "
                       "independent frames per GPU, each encoded then decoded", "frames_per_step": max(1, a.gpus),
           "l2_policy": "inputs larger than L2 (402 MB frame, 805 MB coefficients)"}

    if a.impl == "reference":
        if rank != 0:
            return
        import refharness as R
        if not R.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (needs /root/reference at build time)"}), file=out); out.flush()
            return
        procs = max(1, cores)
        v, ms = run_cpu_reference(procs, max(1, a.steps), max(0, min(a.warmup, 1)))
        sample = "%d processes x one %dx%dx3 12-bit quarter frame per step, in-memory files, ISA level %d" % (
            procs, CPU_TILE, CPU_TILE, R.lib().ojr_cpu_ext_level())
        print(json.dumps({"impl": "reference", "metric": "Mpixels/s encode+decode", "value": v, "unit": "Mpixels/s",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                          "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": "Mpixels/s", "cores": procs, "kind": "reference", "sample": sample},
                          "e2e": {"value": v, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), file=out)
        out.flush()
        return

    import numpy as np
    import torch
    import openjph_b200 as ob
    from openjph_b200 import _lib
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    assert L.ojb_set_device(local) == 0, L.ojb_last_error()
    torch.cuda.set_device(local)
    affinity = bind_to_gpu_numa_node(torch, local)      # before any pinned allocation (first touch)

    p = workload_params()
    frame = make_frame(W, H, 1234 + rank)
    NW = int(os.environ.get("OJB_BENCH_WORKERS", "8"))      # frames in flight per GPU (one codec pair each)
    # pinned host buffers: one input frame (shared, read-only), per-worker outputs
    pin = [torch.empty((H, W), dtype=torch.uint16, pin_memory=True) for _ in range(NC)]
    for t, f in zip(pin, frame):
        t.numpy()[:] = f
    planes = (C.c_void_p * NC)(*[t.data_ptr() for t in pin])
    cs_cap = W * H * NC * 2 + (1 << 20)

    def ck(rc):
        if rc != 0:
            raise RuntimeError(L.ojb_last_error().decode())

    class Worker:
        def __init__(self, params=None):
            self.enc = L.ojb_enc_create(); self.dec = L.ojb_dec_create()
            ck(L.ojb_enc_configure(self.enc, C.byref(params if params is not None else p), ob.U16))
            self.out_pin = [torch.empty((H, W), dtype=torch.uint16, pin_memory=True) for _ in range(NC)]
            self.outs = (C.c_void_p * NC)(*[t.data_ptr() for t in self.out_pin])
            self.cs_pin = torch.empty(cs_cap, dtype=torch.uint8, pin_memory=True)
            self.cs_dev = torch.empty(cs_cap, dtype=torch.uint8, device="cuda")
            self.n = C.c_uint64(); self.fi = _lib.FrameInfo(); self.cs_len = 0

        def e2e(self):
            ck(L.ojb_enc_encode_frame(self.enc, planes, None, self.cs_pin.data_ptr(), cs_cap, C.byref(self.n)))
            ck(L.ojb_dec_read_headers(self.dec, self.cs_pin.data_ptr(), self.n.value, ob.U16, C.byref(self.fi)))
            ck(L.ojb_dec_decode_frame(self.dec, self.outs, None))

        def resident(self):
            ck(L.ojb_enc_encode_resident(self.enc, self.cs_dev.data_ptr(), cs_cap, C.byref(self.n), 1))
            ck(L.ojb_dec_read_headers(self.dec, self.cs_pin.data_ptr(), self.cs_len, ob.U16, C.byref(self.fi)))
            ck(L.ojb_dec_use_device_codestream(self.dec, self.cs_dev.data_ptr()))
            ck(L.ojb_dec_decode_resident(self.dec))

    workers = [Worker() for _ in range(NW)]
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(NW)

    def timed(fn_name, steps, nworkers):
        """`steps` steps of `nworkers` frames each; with several workers every worker streams its own
        `steps` frames back to back (no per-step barrier: frames of one worker overlap the other
        workers' copies and host phases, as in a continuous stream)"""
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if nworkers == 1:
            for _ in range(steps):
                getattr(workers[0], fn_name)()
        else:
            def loop(w):
                f = getattr(w, fn_name)
                for _ in range(steps):
                    f()
            list(pool.map(loop, workers[:nworkers]))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        return dt

    # correctness of what is being timed: lossless round trip through the e2e path, every worker
    for w in workers:
        w.e2e()
        w.cs_len = w.n.value
        for a_, b_ in zip(w.out_pin, frame):
            assert np.array_equal(a_.numpy(), b_), "round trip is not lossless"
        ck(L.ojb_enc_upload_frame(w.enc, planes, None))
    cs_len = workers[0].cs_len
    enc, dec = workers[0].enc, workers[0].dec
    for _ in range(max(3, a.warmup)):
        list(pool.map(lambda w: w.resident(), workers))
    sampler = ClockSampler(local); sampler.start()
    te = (C.c_float * 8)(); td = (C.c_float * 8)()
    # (1) serial pass: one frame at a time -> per-stage CUDA-event times of the kernels
    dt_serial = timed("resident", a.steps, 1)
    L.ojb_enc_timings(enc, te); L.ojb_dec_timings(dec, td)
    # (2) NW frames in flight: host phases and copies of one frame overlap the kernels of another
    dt_res = timed("resident", a.steps, NW)
    for _ in range(max(1, min(a.warmup, 2))):
        list(pool.map(lambda w: w.e2e(), workers))
    dt_e2e = timed("e2e", a.steps, NW)
    te2 = (C.c_float * 8)(); td2 = (C.c_float * 8)()
    L.ojb_enc_timings(enc, te2); L.ojb_dec_timings(dec, td2)
    # secondary workload: the same frame through 9/7 + ICT at Qfactor 90 -- reported in
    # config, not part of `value`
    extra = None
    if os.environ.get("OJB_BENCH_EXTRAS", "1") != "0":
        try:
            pi = ob.make_params(W, H, NC, BD, num_decomps=LEVELS, reversible=False, color_transform=True, qfactor=90)
            keep = workers
            xw = [Worker(pi) for _ in range(min(NW, 4))]
            for w in xw:
                w.e2e(); w.cs_len = w.n.value
                ck(L.ojb_enc_upload_frame(w.enc, planes, None))
            workers = xw
            for _ in range(3):
                list(pool.map(lambda w: w.resident(), xw))
            dtx1 = timed("resident", a.steps, 1)
            tex = (C.c_float * 8)(); tdx = (C.c_float * 8)()
            L.ojb_enc_timings(xw[0].enc, tex); L.ojb_dec_timings(xw[0].dec, tdx)
            dtx = timed("resident", a.steps, len(xw))
            dtxe = timed("e2e", a.steps, len(xw))
            mse = float(np.mean((xw[0].out_pin[1].numpy().astype(np.float64) - frame[1]) ** 2))
            extra = {"workload": "same frame, irreversible 9/7 + ICT, Qfactor 90", "frames_in_flight": len(xw),
                     "Mpixels_per_s": round(W * H * len(xw) * a.gpus * a.steps / dtx / 1e6, 1),
                     "e2e_Mpixels_per_s": round(W * H * len(xw) * a.gpus * a.steps / dtxe / 1e6, 1),
                     "serial_ms_per_frame": round(dtx1 / a.steps * 1e3, 3), "codestream_bytes": int(xw[0].cs_len),
                     "mse_comp1": round(mse, 3),
                     "stages_encode_ms": {k: round(float(v), 4) for k, v in zip(("h2d", "dwt", "ht_encode", "d2h_lengths", "host_wait", "assemble", "d2h_out", "host_ms"), tex)},
                     "stages_decode_ms": {k: round(float(v), 4) for k, v in zip(("h2d", "host_parse", "ht_decode", "dwt_inv", "d2h_image"), tdx)}}
            workers = keep
            for w in xw:
                L.ojb_enc_destroy(w.enc); L.ojb_dec_destroy(w.dec)
        except Exception as e:      # the headline numbers stand on their own
            extra = {"error": str(e)[:200]}
    sampler.stop_flag = True; sampler.join(timeout=2)
    # final gather of the per-rank codestream sizes (the only collective on the path)
    sizes = [cs_len]
    if dist is not None:
        t = torch.tensor([cs_len], device="cuda", dtype=torch.int64)
        g = [torch.zeros_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, g, dst=0)
        if rank == 0:
            sizes = [int(x.item()) for x in g]
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pix = W * H * a.gpus * NW
    value = pix * a.steps / dt_res / 1e6
    e2e = pix * a.steps / dt_e2e / 1e6
    import json as _j
    peaks = {}
    try:
        peaks = _j.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback 6650 GB/s"
    names_e = ("h2d", "dwt", "ht_encode", "d2h_lengths", "host_wait", "assemble", "d2h_out", "host_ms")
    names_d = ("h2d", "host_parse", "ht_decode", "dwt_inv", "d2h_image", "_5", "_6", "host_ms")
    stage_e = {k: round(float(v), 4) for k, v in zip(names_e, te)}
    stage_d = {k: round(float(v), 4) for k, v in zip(names_d, td) if not k.startswith("_")}
    e2e_e = {k: round(float(v), 3) for k, v in zip(names_e, te2)}
    e2e_d = {k: round(float(v), 3) for k, v in zip(names_d, td2) if not k.startswith("_")}
    # dominant kernel = the slowest device stage of the resident step
    samples = W * H * NC
    cand = {"ht_encode": (stage_e["ht_encode"], 4 * samples + cs_len), "ht_decode": (stage_d["ht_decode"], 4 * samples + cs_len),
            "dwt_fwd": (stage_e["dwt"], 2 * samples + 4 * samples * 4 // 3 + 4 * samples // 3),
            "dwt_inv": (stage_d["dwt_inv"], 2 * samples + 4 * samples * 4 // 3 + 4 * samples // 3)}
    dom = max(cand, key=lambda k: cand[k][0])
    ach = cand[dom][1] / (cand[dom][0] * 1e-3) / 1e9 if cand[dom][0] > 0 else 0.0
    # DRAM bytes of the same kernel(s) from the committed ncu --set full capture (profiles/)
    traffic = None
    try:
        prof = _j.load(open(os.path.join(ROOT, "profiles", PROFILE_JSON)))
        keys = {"ht_encode": ["ht_encode_serial", "ht_encode"], "ht_decode": ["ht_decode_serial", "ht_dec_fill", "ht_dec_step1", "ht_dec_step2"],
                "dwt_fwd": ["dwt_fwd"], "dwt_inv": ["dwt_inv"]}[dom]
        traffic = sum(int(prof[k]["traffic_bytes"]) for k in keys if k in prof)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
            "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": cand[dom][1], "ms_per_launch": cand[dom][0],
            "note": "entropy-coding kernels are instruction-issue / latency bound (thread-per-block, ncu: 52-56 % "
                    "issue-active at 16 % occupancy, DRAM traffic ~= algorithmic bytes); see profiles/README.md"}
    # A = S_in + S_out per frame (and the two-pass budget A2 = A + 2*4*W*H*C), per direction
    A = 2 * samples + cs_len
    A2 = A + 8 * samples
    t_enc = (stage_e["dwt"] + stage_e["ht_encode"] + stage_e["assemble"]) * 1e-3
    t_dec = (stage_d["ht_decode"] + stage_d["dwt_inv"]) * 1e-3
    roof["frame"] = {"A_bytes": A, "A2_bytes": A2,
                     "encode_kernels_ms": round(t_enc * 1e3, 3), "decode_kernels_ms": round(t_dec * 1e3, 3),
                     "encode_frac_A": round(A / t_enc / 1e9 / peak, 4), "encode_frac_A2": round(A2 / t_enc / 1e9 / peak, 4),
                     "decode_frac_A": round(A / t_dec / 1e9 / peak, 4), "decode_frac_A2": round(A2 / t_dec / 1e9 / peak, 4),
                     "encode_Mpix_s_device": round(W * H / t_enc / 1e6, 1), "decode_Mpix_s_device": round(W * H / t_dec / 1e6, 1)}
    cfg.update({"frames_per_step": NW * a.gpus, "frames_in_flight_per_gpu": NW, "host_affinity": affinity, "irv97_ict_q90": extra, "e2e_stages_encode_ms": e2e_e, "e2e_stages_decode_ms": e2e_d,
                "serial_ms_per_frame": round(dt_serial / a.steps * 1e3, 3),
                "serial_Mpixels_per_s": round(W * H * a.steps / dt_serial / 1e6, 1),
                "stages_encode_ms": stage_e, "stages_decode_ms": stage_d, "codestream_bytes": sizes,
                "pipeline_hbm_frac": round(((2 * samples + cs_len) * 2 * NW) / (dt_res / a.steps) / 1e9 / peak, 4)})
    res = {"metric": "Mpixels/s encode+decode", "value": value, "unit": "Mpixels/s", "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": dt_res / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg,
           "clocks": sampler.summary(),
           "e2e": {"value": e2e, "unit": "Mpixels/s", "h2d_bytes_per_step": (W * H * NC * 2 + cs_len) * a.gpus * NW,
                   "d2h_bytes_per_step": (W * H * NC * 2 + cs_len) * a.gpus * NW, "ms_per_step": dt_e2e / a.steps * 1e3},
           "gpu_launches": int(L.ojb_enc_kernel_launches(enc) + L.ojb_dec_kernel_launches(dec)) * a.steps * NW,
           "roofline": roof}
    if not a.no_cpu_baseline and world == 1 or (rank == 0 and not a.no_cpu_baseline):
        import refharness as R
        if R.available():
            cpu_worker_init(1234)
            te_ = td_ = 0.0
            for _ in range(2):
                e_, d_, _n = cpu_worker((1234, False)); te_ += e_; td_ += d_
            v1 = 2 * CPU_TILE * CPU_TILE / (te_ + td_) / 1e6
            res["cpu_baseline"] = {"value": v1, "unit": "Mpixels/s", "cores": 1, "kind": "reference",
                                   "sample": "2 x one %dx%dx3 12-bit quarter frame, 1 thread (the library's native mode), "
                                             "in-memory; encode %.1f Mpix/s decode %.1f Mpix/s; ISA level %d" % (
                                                 CPU_TILE, CPU_TILE, 2 * CPU_TILE * CPU_TILE / te_ / 1e6,
                                                 2 * CPU_TILE * CPU_TILE / td_ / 1e6, R.lib().ojr_cpu_ext_level())}
    print(json.dumps(res), file=out); out.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
