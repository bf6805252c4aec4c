#!/usr/bin/env python
"""bench.py - decoded 15-s FT8 slots/s on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--slots S] [--impl reference]

Workload (config.workload): BASELINE.json configs[2] - a batch of 4096 independent synthetic 15-s
slots per GPU (12 kHz, one FT8 signal each, SNR -16..-4 dB, float32 audio), decoded with
bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20.
A step = one pass of the whole hot path (STFT -> Costas sync -> selection -> LLR -> LDPC/CRC) over the
batch.  `value` times it with the audio already resident in HBM; `e2e` times the public API
(ft8_decode.decode_slots) from pinned HOST memory to HOST results, copies inside the timed region.
The batch (2.95 GB) is far larger than L2 (126 MB), so no flush is needed between steps.

N > 1: launched by torchrun, one rank per GPU; slots shard across ranks with no data-path collective
(weak scaling: 4096 slots per GPU); time = max over ranks.

--impl reference: the reference's CPU path for the same metric.  The reference is pure Python and does
not travel to the GPU box, so this arm runs the CPU oracle port (oracle/ft8_oracle.c, a compiled C
restatement - far faster than the reference's Python loops) on all host threads, on a bounded sample
of the same workload.
"""
import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# keep stdout to the one JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

FS, N_SAMPLES = 12000, 180000
DECODE_KW = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
WORKLOAD = ("C3: batch of %d independent synthetic 15-s FT8 slots per GPU, 12 kHz float32, one signal per slot at "
            "SNR -16..-4 dB, osr 2x2, K=20, min_score=1, max_iterations=20")
# algorithmic bytes per unit (float64 spectrogram as computed here; stated in DESIGN.md)
STFT_BYTES_PER_SLOT = 180000 * 4 + 1920 * 186 * 8      # read audio f32 + write dB f64
SYNC_BYTES_PER_SLOT = 1920 * 186 * 8 + 88 * 1906 * 8   # read dB f64 + write score map f64
# the selection reads the head of the score map until K passers are found (one 4096-score pass here: min_score = 1 is
# passed at once) plus the 60 strip arg-maxima the sync kernel leaves; it no longer re-reads the map
SELECT_BYTES_PER_SLOT = 4096 * 8 + 60 * 16
LDPC_BYTES_PER_WORD = 174 * 8 + 10 + 1 + 3 * 4
# SURVEY.md 8d quotes the same stages for a float32 spectrogram (a design this repo rejects);
# reported next to the float64 figures so both readings of the contract are on the line
SURVEY_FP32_BYTES = {"stft_db": 180000 * 4 + 1920 * 186 * 4, "costas_sync": 1920 * 186 * 4 + 20 * 12 + 4}
STAGES = ["stft_db", "costas_sync", "select", "llr", "ldpc_crc"]


def workload_config(slots, chunk):
    """`config` of the JSON line - the same object on both arms (the reference arm runs a bounded sample of this
    workload; the sample is described in its cpu_baseline)."""
    return {"workload": WORKLOAD % slots, "slots_per_gpu": slots, "l2_policy": "inputs (2.95 GB/GPU) larger than L2, no flush",
            "chunk_slots": chunk or "library default", "sharding": "slots split across ranks, no collective"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--slots", type=int, default=4096, help="slots per GPU")
    ap.add_argument("--chunk", type=int, default=0, help="slots per L2 chunk (0 = library default)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="slots in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """SM clock, power and throttle reasons sampled WHILE the timed region runs: NVML every 5 ms from a thread
    (the timed region of the default run lasts ~0.1 s), `nvidia-smi -lms 200` if NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # NVML clocks-event (throttle) reason bits
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
            "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.handle, self.samples, self.stop_flag, self.thread = None, None, [], False, None
        self.recording = True          # the polling thread runs from before the warm-up; samples count only while this is set
        self.errors = 0

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            pr = torch.cuda.get_device_properties(self.idx)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
        return pynvml, h

    def _poll(self):
        n, h = self.nvml, self.handle
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                smp = (n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx, n.nvmlDeviceGetPowerUsage(h) / 1000.0, rs)
                if self.recording:
                    self.samples.append(smp)
            except Exception:
                self.errors += 1
            time.sleep(0.005)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            sm = [x[0] for x in self.samples]
            reasons = sorted(nm for nm, bit in self.BITS.items() if any(x[3] & bit for x in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.samples[0][1]) if sm else None,
                    "power_w_max": max(x[2] for x in self.samples) if sm else None, "samples": len(sm), "reasons": reasons,
                    "nvml_errors": self.errors, "source": "nvml, 5 ms period, during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 200"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic_per_slot(stage):
    """DRAM bytes per slot of a stage from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(p))[stage]["dram_bytes_per_slot"])
    except Exception:
        return None


def measured_pipe_utilisation(stage):
    """FP64-pipe / issue-slot / DRAM utilisation of a stage's kernel from the same committed ncu capture (a static
    figure from profiles/, not a live measurement): what actually bounds these float64 kernels."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))[stage]
        return {"fp64_pipe_pct": d["fp64_pipe_pct"], "issue_active_pct": d["issue_active_pct"], "dram_pct": d["dram_pct"],
                "lsu_pipe_pct": d.get("lsu_pipe_pct")}
    except Exception:
        return None


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_port(waves_f32: np.ndarray, threads: int):
    """Oracle port over `waves`, all threads.  Returns (seconds, per-slot record lists)."""
    import ft8_oracle as orc
    orc.lib()
    x = waves_f32.astype(np.float64)
    t0 = time.perf_counter()
    recs = orc.decode_batch(x, FS, DECODE_KW["bins_per_tone"], DECODE_KW["steps_per_symbol"], DECODE_KW["max_candidates"],
                            DECODE_KW["min_score"], DECODE_KW["max_iterations"], n_threads=threads)
    return time.perf_counter() - t0, recs


def reference_arm(args):
    """--impl reference: CPU oracle port on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ft8_demodulator_b200 import synth
    threads = cpu_threads()
    sample = args.cpu_sample or max(16, min(512, 32 * threads))      # ~3 s of wall clock per step on 16 threads
    waves = np.stack([synth.c3_slot(i)[0] for i in range(sample)]).astype(np.float32)
    for _ in range(args.warmup):
        run_cpu_port(waves[:max(1, min(sample, threads))], threads)
    t = 0.0
    for _ in range(args.steps):
        dt, _ = run_cpu_port(waves, threads)
        t += dt
    val = sample * args.steps / t
    line = {
        "impl": "reference", "metric": "decoded 15-s FT8 slots/sec", "value": val, "unit": "slots/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.slots, args.chunk),
        "cpu_baseline": {"value": val, "unit": "slots/s", "cores": threads, "kind": "port",
                         "sample_slots_per_step": sample,
                         "sample": "%d C3 slots per step through oracle/ft8_oracle.c (compiled C restatement of the "
                                   "reference's Python path; the Python reference itself measured 0.064 slots/s per core "
                                   "in the build container, BASELINE.md section 2)" % sample},
        "e2e": {"value": val, "unit": "slots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from ft8_demodulator_b200 import _lib, dist as fdist, engine, synth
    from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd

    rank, local_rank, world = fdist.init_process_group()
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    numa = fdist.bind_host_to_gpu_numa(local_rank)      # before any pinned buffer exists
    dev = torch.device("cuda", local_rank)
    L = _lib.lib()
    B = args.slots

    # ---- synthetic C3 batch, generated on the device (untimed) ----
    waves, payloads = synth.synth_batch_torch(B, dev, seed=1234 + rank)
    g = fd.decode_geometry(N_SAMPLES, FS, DECODE_KW["bins_per_tone"], DECODE_KW["steps_per_symbol"])
    params = engine.make_params(FS, DECODE_KW["bins_per_tone"], DECODE_KW["steps_per_symbol"], DECODE_KW["max_candidates"],
                                DECODE_KW["min_score"], DECODE_KW["max_iterations"], g["bin_lo"], g["bin_hi"],
                                g["frame_lo"], g["frame_hi"], args.chunk)
    dec = engine.SlotDecoder(B, N_SAMPLES, params, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler's thread starts here, so NVML's first-call latency is not paid inside the ~0.1 s timed
    # region; its samples only count while `recording` is set) ----
    sampler = ClockSampler(local_rank)
    sampler.recording = False
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        dec.run(waves)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident input ----
    L.ft8b_stage_timing(1)
    launches0 = L.ft8b_launch_count()
    barrier()                          # (the first NCCL barrier of a multi-rank run takes seconds: not part of the clock samples)
    sampler.recording = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        dec.run(waves)
    e1.record()
    torch.cuda.synchronize()
    sampler.recording = False
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = L.ft8b_launch_count() - launches0
    stage_ms = (C.c_double * 5)()
    stage_n = (C.c_int64 * 5)()
    L.ft8b_stage_times(stage_ms, stage_n)
    L.ft8b_stage_timing(0)
    n_in_region = len(sampler.samples)
    if rank == 0 and n_in_region < 3:
        # the region is short (K x 16 ms): when NVML answered fewer than three times inside it, keep the SAME load running
        # (untimed repeats of the step) until there are samples to judge the clocks by
        sampler.recording = True
        t_top = time.time()
        while len(sampler.samples) < 8 and time.time() - t_top < 3.0:
            dec.run(waves)
            torch.cuda.synchronize()
        sampler.recording = False
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["samples_in_timed_region"] = n_in_region
        if n_in_region < 3:
            clocks["source"] = clocks.get("source", "") + "; topped up under the same load right after it (untimed repeats of the step)"
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- decode quality / round-trip property on the full batch ----
    out = dec.out.cpu()
    ok = out.ok.numpy().astype(bool)
    pl = out.payload.numpy()
    hit = sum(1 for i in range(B) if any(bytes(pl[i, c]) == bytes(payloads[i]) for c in np.nonzero(ok[i])[0]))
    n_words = int(out.n_cand.sum())

    # ---- e2e: public API, pinned host input -> host results ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty((B, N_SAMPLES), dtype=torch.float32, pin_memory=True)
        host.copy_(waves)
        torch.cuda.synchronize()
        kw = dict(DECODE_KW)
        res = None
        for _ in range(2):
            res = fd.decode_slots(host, FS, chunk_slots=args.chunk, **kw)
        barrier()
        gc.collect()
        gc.disable()        # a generational collection inside three 55-ms host calls reads as a 10 % outlier
        t0 = time.perf_counter()
        dt_dev = 0.0
        e2e_steps = max(2, min(args.steps, 3))
        for _ in range(e2e_steps):
            res = fd.decode_slots(host, FS, chunk_slots=args.chunk, **kw)
            dt_dev += engine._lib.lib().ft8b_last_host_call_ms() * 1e-3
        dt = time.perf_counter() - t0
        gc.enable()

        def timed(fn, n=2):
            keep = fn()      # two warm calls whose results stay alive: the page-locked result arrays of large batches come
            fn()             # from torch's caching host allocator, and both sets in rotation must exist before timing
            del keep
            barrier()
            t = time.perf_counter()
            for _ in range(n):
                r = fn()
            return (time.perf_counter() - t) / n, r

        # the same batch (a) from pageable NumPy memory (staged through pinned buffers by the library's host threads) and
        # (b) through the earlier torch-stream version of the overlap (device results, then .cpu()): information only
        host_np = host.numpy().copy()
        dt_np, res_np = timed(lambda: fd.decode_slots(host_np, FS, chunk_slots=args.chunk, **kw))
        dt_th, res_th = timed(lambda: fd.decode_slots(host, FS, chunk_slots=args.chunk, return_device=True, **kw).cpu())
        del host_np
        if world > 1:
            t = torch.tensor([dt, dt_dev, dt_np, dt_th], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_dev, dt_np, dt_th = [float(x) for x in t.tolist()]
        same = all(torch.equal(getattr(res.arrays, f), getattr(out, f)) for f in out.__dataclass_fields__)
        same_np = all(torch.equal(getattr(res_np.arrays, f), getattr(out, f)) for f in out.__dataclass_fields__)
        same_th = all(torch.equal(getattr(res_th, f), getattr(out, f)) for f in out.__dataclass_fields__)
        e2e = {"value": world * B * e2e_steps / dt, "unit": "slots/s", "h2d_bytes_per_step": int(host.numel() * 4),
               "d2h_bytes_per_step": int(res.arrays.nbytes()), "ms_per_step": 1e3 * dt / e2e_steps,
               "call": "decode_slots(pinned host float32) -> ft8b_decode_slots_host (C ABI, host pointers in and out)",
               "timing": "host wall clock around the calls (results are host arrays when a call returns), max over ranks",
               "cuda_event_ms_per_step": 1e3 * dt_dev / e2e_steps,
               "identical_to_device_resident_run": bool(same),
               "pageable_numpy_input": {"value": world * B / dt_np, "unit": "slots/s", "identical": bool(same_np)},
               "torch_stream_overlap": {"value": world * B / dt_th, "unit": "slots/s", "identical": bool(same_th)}}
        # the link itself: the same pinned batch copied host-to-device with nothing else going on (what bounds e2e)
        dst = torch.empty_like(waves)
        dst.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(2):
            dst.copy_(host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link_gbs = 2 * host.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del dst
        e2e["h2d_link"] = {"plain_copy_gbs": link_gbs, "e2e_gbs": host.numel() * 4 * e2e_steps / dt / 1e9,
                           "e2e_frac_of_plain_copy": (host.numel() * 4 * e2e_steps / dt / 1e9) / link_gbs,
                           "note": "e2e moves 720 kB per slot over PCIe; a plain cudaMemcpyAsync of the same pinned batch is the ceiling"}
        del host
        # same call fed 16-bit PCM (what a WAV file / SDR capture holds; from_wave.py semantics): half the PCIe bytes.
        # Extra information only - the audio is a 16-bit quantisation of the batch, so results are not comparable
        # slot by slot with the float32 run; they are checked against the float32 path fed the same quantised values.
        peak = waves.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
        q16 = torch.round(waves / peak * 32767.0).to(torch.int16)
        host16 = torch.empty((B, N_SAMPLES), dtype=torch.int16, pin_memory=True)
        host16.copy_(q16)
        torch.cuda.synchronize()
        for _ in range(2):
            r16 = fd.decode_slots(host16, FS, chunk_slots=args.chunk, pcm16=True, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            r16 = fd.decode_slots(host16, FS, chunk_slots=args.chunk, pcm16=True, **kw)
        torch.cuda.synchronize()
        dt16 = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt16], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt16 = float(t.item())
        nchk = min(B, 256)
        # NumPy float32 division on the host = from_wave.py's normalisation (torch's tensor/scalar multiplies by 1/b)
        ref16 = fd.decode_slots(q16[:nchk].cpu().numpy().astype(np.float32) / np.float32(32767), FS,
                                chunk_slots=args.chunk, **kw)
        same16 = all(torch.equal(getattr(r16.arrays, f)[:nchk], getattr(ref16.arrays, f)) for f in out.__dataclass_fields__)
        e2e["pcm16"] = {"value": world * B * e2e_steps / dt16, "unit": "slots/s", "h2d_bytes_per_step": int(host16.numel() * 2),
                        "ms_per_step": 1e3 * dt16 / e2e_steps, "identical_to_float32_path_on_same_samples": bool(same16)}
        del host16, q16

    # ---- optional gather of the fixed-size decode records (the only collective; untimed above) ----
    gather_ms = None
    if world > 1:
        rec = fdist.pack_records(dec.out)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        allrec = fdist.gather_records(rec, world * B, rank, world)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        assert allrec.shape[0] == world * B

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- CPU baseline (oracle port) on a bounded sample of the same batch + parity of that sample ----
    cpu = None
    parity = None
    if not args.no_cpu and world == 1:      # the CPU leg runs at N=1 only (rank 0)
        threads = cpu_threads()
        sample = args.cpu_sample or max(16, min(4096, 256 * threads))    # the whole batch on 16 threads: ~24 s of wall clock
        sample = min(sample, B)
        wave_s = waves[:sample].cpu().numpy()
        secs, recs = run_cpu_port(wave_s, threads)
        cpu = {"value": sample / secs, "unit": "slots/s", "cores": threads, "kind": "port",
               "sample": "first %d slots of the batch through oracle/ft8_oracle.c on %d threads (%.1f s); the Python "
                         "reference itself: 0.064 slots/s per core (build container, BASELINE.md section 2)" % (sample, threads, secs)}
        bad = 0
        for i in range(sample):
            r = recs[i]
            same = int(out.n_cand[i]) == len(r)
            for c, rr in enumerate(r if same else []):
                same &= (int(out.cand_time[i, c]), int(out.cand_freq[i, c])) == (rr["abs_time"], rr["abs_freq"])
                same &= bool(out.ok[i, c]) == rr["ok"] and int(out.ldpc_errors[i, c]) == rr["ldpc_errors"]
                same &= bytes(pl[i, c]) == rr["payload"]
            bad += 0 if same else 1
        parity = {"slots_checked_vs_oracle": sample, "slots_identical": sample - bad}

    # ---- roofline of the dominant kernel (largest share of the timed region) ----
    peak, peak_src = measured_peak_gbs()
    sm = [float(stage_ms[i]) for i in range(5)]
    sn = [int(stage_n[i]) for i in range(5)]
    units = [B * args.steps, B * args.steps, B * args.steps, n_words * args.steps, n_words * args.steps]
    alg = [STFT_BYTES_PER_SLOT, SYNC_BYTES_PER_SLOT, SELECT_BYTES_PER_SLOT, 58 * 8 * 8 + 174 * 8, LDPC_BYTES_PER_WORD]
    stages = {}
    for i, nm in enumerate(STAGES):
        if sm[i] > 0:
            gbs = alg[i] * units[i] / (sm[i] * 1e-3) / 1e9
            stages[nm] = {"ms_per_step": sm[i] / args.steps, "launches_per_step": sn[i] / args.steps,
                          "share": sm[i] / sum(sm), "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    dom = max(range(5), key=lambda i: sm[i])
    tps = measured_traffic_per_slot(STAGES[dom])
    slots_per_launch = units[dom] / max(sn[dom], 1)
    roof = {"kernel": STAGES[dom], "bound": "hbm", "achieved": stages[STAGES[dom]]["achieved_gbs"], "peak": peak,
            "unit": "GB/s", "frac": stages[STAGES[dom]]["achieved_gbs"] / peak,
            "traffic": (tps * slots_per_launch) if (tps and dom < 3) else None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg[dom] * units[dom] / max(sn[dom], 1),
            "algorithmic_bytes_per_slot": alg[dom] if dom < 3 else None,
            "survey_fp32_layout": ({"bytes_per_slot": SURVEY_FP32_BYTES[STAGES[dom]],
                                    "achieved": SURVEY_FP32_BYTES[STAGES[dom]] * units[dom] / (sm[dom] * 1e-3) / 1e9,
                                    "frac": SURVEY_FP32_BYTES[STAGES[dom]] * units[dom] / (sm[dom] * 1e-3) / 1e9 / peak}
                                   if STAGES[dom] in SURVEY_FP32_BYTES else None),
            "avg_launch_ms": sm[dom] / max(sn[dom], 1),
            "ncu_utilisation": measured_pipe_utilisation(STAGES[dom]),
            "note": "float64 FFT/stencil/BP kernels are bound by the shared-memory (LSU) data pipe and the FP64 pipe, not HBM (DESIGN.md)"}

    line = {
        "metric": "decoded 15-s FT8 slots/sec", "value": value, "unit": "slots/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(B, args.chunk),
        "candidates_per_s": value * 88 * 1906, "ldpc_codewords_per_s": world * n_words / (ms_step * 1e-3),
        "roofline": roof, "stages": stages, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks, "decode_check": {"slots_with_tx_payload_decoded": hit, "slots": B}, "parity": parity,
        "gather_ms": gather_ms, "host_numa_binding": numa,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
