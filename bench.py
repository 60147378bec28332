#!/usr/bin/env python3
"""bench.py -- MP2 DAB encode throughput (audio-seconds encoded per second) on N B200s.

Workload (BASELINE.json configs[1]): MP2 DAB 192 kbit/s, 48 kHz stereo, psy model 1, a 10 h synthetic PCM batch per
GPU.  One step = one pass of the whole encode path over that batch.  Ranks hold independent streams (no
collective on the data path; weak scaling), timing is max over ranks between barriers.

  value      device-resident: PCM already in HBM, frames written to HBM      (tlb_batch_encode_device)
  e2e        through the C ABI with pinned HOST buffers, H2D/D2H inside      (tlb_batch_encode)
  roofline   dominant kernel, CUDA-event time measured in the timed region on the launching stream
  cpu_baseline  the reference libtoolame-dab (oracle/_ref/ref_driver, compiled unmodified) one process per host
             core on a bounded sample of the same signal; falls back to the oracle port when _ref is absent

`--impl reference` runs only that CPU arm (rank 0) and prints it in the same JSON shape.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json configs: B = configs[1] (the metric's configuration, the default), C = configs[2], E = configs[4]
CONFIGS = {"B": (48000, "j", 192, 2, 1, "MP2 DAB 192 kbit/s 48 kHz stereo (mode j), psy model 1"),
           "C": (24000, "m", 64, 1, 1, "MP2 DAB 64 kbit/s 24 kHz mono (LSF tables), psy model 1"),
           "E": (48000, "j", 256, 2, 2, "MP2 DAB 256 kbit/s 48 kHz joint stereo, psy model 2")}
FS, MODE, KBPS, NCH, PSY, WORKLOAD = CONFIGS["B"]   # odr-audioenc's default mode for 2 channels is joint stereo
METRIC = "MP2 audio-seconds encoded/sec"
UNIT = "audio-s/s"
KERNEL_ALG = {}


def select_config(name):
    """algorithmic bytes and FP64 flops per frame (data-independent flops), per kernel: the bytes
    a kernel must move for one frame and the FP64 operations it must do."""
    global FS, MODE, KBPS, NCH, PSY, WORKLOAD, KERNEL_ALG
    FS, MODE, KBPS, NCH, PSY, WORKLOAD = CONFIGS[name]
    lg = (3 if FS == 48000 else 6) * KBPS
    sbl = 27 if FS == 48000 else 30
    KERNEL_ALG = {
        "k_filterbank": (NCH * 1152 * 2 + NCH * 1152 * 8 + 192 + 96, NCH * 74844 + (2304 if MODE == "j" else 0)),
        "k_spectrum": (NCH * (1024 * 2 + 2 * 512 * 8 + 128 + 256), NCH * 27334),
        "k_label": (NCH * (2 * 512 * 8 + 128 + 1128), 0),
        "k_threshold": (NCH * (1128 + 256 + 256) + 192, 0),
        "k_spectrum2": (NCH * 2 * (1024 * 2 + 2 * 513 * 8), NCH * 2 * (1024 + 20488 + 2046)),
        "k_psy2": (NCH * (4 * 2 * 513 * 8 + 256), NCH * 2 * 513 * 40),
        "k_alloc": (192 + NCH * 32 * 8 + 336, 0),
        "k_pack": (NCH * 1152 * 8 + 336 + 96 + lg, NCH * sbl * 36 * 4),
    }


select_config("B")

# dram__bytes_read.sum + dram__bytes_write.sum per frame of config B, from the ncu --set full capture summarised in
# profiles/ncu_r1_summary.md (end of round, launches of 75 777 frames); bench.py scales it to its own launch size
NCU_DRAM_BYTES_PER_FRAME = {"k_filterbank": 22774, "k_spectrum": 20479, "k_label": 30285, "k_threshold": 1966,
                            "k_alloc": 623, "k_pack": 19667}


def synth_pcm_torch(n_frames, seed, device):
    """S1-style signal of SURVEY.md 8(d) generated on the GPU in one-minute pieces (tones under a slow envelope,
    a wandering 3 kHz component and noise); int16 (n_samples, 2)."""
    import torch
    n = n_frames * 1152
    out = torch.empty((n, NCH), dtype=torch.int16, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    step = FS * 60
    c = torch.arange(NCH, dtype=torch.float64, device=device)[None, :]
    for s0 in range(0, n, step):
        m = min(step, n - s0)
        t = (torch.arange(s0, s0 + m, dtype=torch.float64, device=device) / FS)[:, None]
        noise = torch.rand((m, NCH), generator=g, dtype=torch.float64, device=device) * 2 - 1
        env = 0.5 + 0.5 * torch.sin(2 * torch.pi * 0.37 * t)
        v = env * (0.3 * torch.sin(2 * torch.pi * (440 + 110 * c) * t)
                   + 0.2 * torch.sin(2 * torch.pi * (3000 + 500 * torch.sin(t)) * t)) + 0.05 * noise
        out[s0:s0 + m] = torch.round(v * 32767 * 0.8).to(torch.int16)
    return out


def bind_to_gpu_numa_node(index):
    """Run this rank (and allocate its pinned host buffers) on the CPUs next to its GPU: with several ranks per
    node the host<->device copies otherwise cross the socket interconnect.  Best effort; returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and 64 * w + b < n_cpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "cpus %d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))
    except Exception as e:  # noqa: BLE001
        return "unbound (%s)" % type(e).__name__
    return "unbound"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                  "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self._stop_evt.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([x.strip() for x in line.split(",")])
        p.terminate()

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(len(r) > i and r[i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        mx = max(int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit())
        # median of the upper half = clocks under load (the sampler also sees idle gaps between steps)
        return {"sm_mhz": sm[(len(sm) * 3) // 4], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(seconds_audio, n_procs, steps=1, warmup=0):
    """The reference's own CPU implementation on a bounded sample: n_procs processes (one stream each, the
    reference is not re-entrant) encode `seconds_audio` of the S1 signal concurrently.  Returns (per-step wall
    seconds list, kind, frames per process)."""
    import numpy as np
    import signals
    n_frames = int(seconds_audio * FS) // 1152
    pcm = signals.make("S1", n_frames, NCH, FS)
    ref_driver = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    times = []
    with tempfile.TemporaryDirectory() as td:
        pin = os.path.join(td, "in.pcm")
        pcm.tofile(pin)
        if os.path.exists(ref_driver):
            kind = "reference"
            cmd = [ref_driver, str(FS), MODE, str(KBPS), str(PSY), "0", pin, os.path.join(td, "out%d.mp2"), "--bench"]

            def one_step():
                t0 = time.perf_counter()
                ps = [subprocess.Popen([c if "%d" not in c else c % i for c in cmd], stdout=subprocess.PIPE,
                                       stderr=subprocess.DEVNULL, text=True) for i in range(n_procs)]
                enc = [json.loads(p.communicate()[0])["seconds"] for p in ps]
                wall = time.perf_counter() - t0
                return max(enc), wall  # encode-loop time only (PCM preloaded), slowest process
        else:
            kind = "port"
            import multiprocessing as mp

            def one_step():
                with mp.get_context("fork").Pool(n_procs) as pool:
                    enc = pool.map(_oracle_worker, [pin] * n_procs)
                return max(enc), max(enc)
        for i in range(warmup + steps):
            enc_s, _ = one_step()
            if i >= warmup:
                times.append(enc_s)
    return times, kind, n_frames


def _oracle_worker(pin):
    import numpy as np
    import oracle
    pcm = np.fromfile(pin, dtype=np.int16).reshape(-1, NCH)
    c = oracle.configure(FS, MODE, KBPS, PSY)
    t0 = time.perf_counter()
    oracle.encode(c, pcm)
    return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_s = args.cpu_sample_seconds
    times, kind, n_frames = cpu_reference_run(sample_s, cores, steps=args.steps, warmup=min(args.warmup, 1))
    audio_s = n_frames * 1152 / FS * cores
    t = sum(times) / len(times)
    value = audio_s / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD + "; per step a bounded sample of the "
                               "10 h batch: %.0f s of audio per host core, %d cores" % (n_frames * 1152 / FS, cores)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d processes x %d frames (%.0f s) of signal S1, encode loop only" % (cores, n_frames, n_frames * 1152 / FS)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


ENSEMBLE = [(48000, "j", 192)] * 6 + [(48000, "j", 160)] * 4 + [(48000, "j", 128)] * 4 + [(48000, "j", 112)] * 2 + \
           [(48000, "m", 96)] * 2   # BASELINE configs[3]: 18 MP2 services, mixed 96-192 kbit/s


def run_ensemble(args):
    """--config D: the 18-service ensemble, `--hours` of audio per service, sharded by whole services across the
    ranks (odr_audioenc_b200.sharding.service_shards: LPT by a bitrate-weighted cost).  Fixed total work: strong
    scaling.  Device-resident value and end-to-end (pinned host buffers) like the main mode."""
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    import odr_audioenc_b200 as tl
    from odr_audioenc_b200 import sharding
    global FS, NCH
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_frames = int(round(args.hours * 3600 * 48000 / 1152))
    plan = sharding.service_shards([(fs, 1 if m == "m" else 2, br, n_frames) for fs, m, br in ENSEMBLE], world)
    mine = plan[rank]
    L = tl.lib()
    encs, work = {}, []
    for i in mine:
        fs, mode, br = ENSEMBLE[i]
        key = (fs, mode, br)
        if key not in encs:
            encs[key] = tl.BatchEncoder(fs, mode, br, 1, 0, local, args.chunk_frames or 148 * 256)
        e = encs[key]
        FS, NCH = fs, e.nch
        d_pcm = synth_pcm_torch(n_frames, 5000 + i, dev)
        d_out = torch.empty(n_frames * e.lg_frame, dtype=torch.uint8, device=dev)
        work.append((e, d_pcm, d_out))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        for e, d_pcm, d_out in work:
            e.encode_device(d_pcm.data_ptr(), n_frames, 0, False, None, d_out.data_ptr())
        for e in encs.values():
            e.sync()

    h_bufs = []
    for e, d_pcm, d_out in work:
        hp, ho = L.tlb_host_alloc(d_pcm.numel() * 2), L.tlb_host_alloc(d_out.numel())
        h_pcm = np.ctypeslib.as_array((C.c_int16 * d_pcm.numel()).from_address(hp)).reshape(-1, e.nch)
        h_out = np.ctypeslib.as_array((C.c_uint8 * d_out.numel()).from_address(ho))
        torch.from_numpy(h_pcm).copy_(d_pcm)
        h_bufs.append((h_pcm, h_out))
    L.tlb_batch_encode_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]

    def step_host():
        for (e, _, _), (h_pcm, h_out) in zip(work, h_bufs):
            L.tlb_batch_encode_async(e._h, h_pcm.ctypes.data, n_frames, 0, 0, None, h_out.ctypes.data)
        for e in encs.values():
            e.sync()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        barrier()
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t

    launches0 = sum(e.launches for e in encs.values())
    t_dev = timed(step_device)
    launches = (sum(e.launches for e in encs.values()) - launches0) // (args.steps + args.warmup)
    t_host = timed(step_host)
    same = all(bool(torch.equal(torch.from_numpy(h_out[:4096]).to(dev), d_out[:4096])) for (_, _, d_out), (_, h_out) in zip(work, h_bufs))
    if rank == 0:
        audio = len(ENSEMBLE) * n_frames * 1152 / 48000
        print(json.dumps({
            "metric": METRIC, "value": audio * args.steps / t_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "DAB ensemble: 18 MP2 services x %.3g h, 6x192 4x160 4x128 2x112 kbit/s joint stereo + 2x96 mono, "
                                   "psy model 1, whole services per GPU (LPT)" % args.hours, "plan": plan},
            "e2e": {"value": audio * args.steps / t_host, "unit": UNIT,
                    "h2d_bytes_per_step": sum(int(n_frames * 1152 * (1 if m == "m" else 2) * 2) for _, m, _ in ENSEMBLE),
                    "d2h_bytes_per_step": sum(int(n_frames * 3 * br) for _, _, br in ENSEMBLE)},
            "gpu_launches": int(launches * args.steps), "host_equals_device_outputs": same}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hours", type=float, default=10.0, help="audio per GPU per step (BASELINE config: 10 h)")
    ap.add_argument("--config", default="B", choices=sorted(CONFIGS) + ["D"], help="B = BASELINE configs[1] (default), C = configs[2], E = configs[4]")
    ap.add_argument("--chunk-frames", type=int, default=0)
    ap.add_argument("--cpu-sample-seconds", type=float, default=120.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 0)
    if args.config == "D":
        if args.impl == "reference":
            raise SystemExit("--config D has no reference arm (use the default config)")
        if args.hours == 10.0:
            args.hours = 1.0  # BASELINE configs[3]: 1 h per service
        return run_ensemble(args)
    select_config(args.config)
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import odr_audioenc_b200 as tl
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this implementation has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_frames = int(round(args.hours * 3600 * FS / 1152))
    enc = tl.BatchEncoder(FS, MODE, KBPS, PSY, 0, local, args.chunk_frames)
    lg = enc.lg_frame
    L = tl.lib()

    # ---- inputs: resident in HBM (value) and in pinned host memory (e2e)
    d_pcm = synth_pcm_torch(n_frames, 1000 + rank, dev)
    d_out = torch.empty(n_frames * lg, dtype=torch.uint8, device=dev)
    pcm_bytes, out_bytes = d_pcm.numel() * 2, d_out.numel()

    def step_device():
        enc.encode_device(d_pcm.data_ptr(), n_frames, 0, False, None, d_out.data_ptr())

    stream = torch.cuda.ExternalStream(enc.stream, device=dev)

    def timed(fn, steps, warmup, sync_each):
        """max-over-ranks seconds for `steps` steps, device-timed with CUDA events on the encoder's stream"""
        for _ in range(warmup):
            fn()
            enc.sync()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
            if sync_each:
                enc.sync()
        e1.record(stream)
        enc.sync()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) * 1e-3
        barrier()
        return dev_s, wall

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = enc.launches
    dev_s, _ = timed(step_device, args.steps, args.warmup, sync_each=False)
    launches_per_step = (enc.launches - launches0) // (args.steps + args.warmup)
    # second timed region, same work: CUDA events around every kernel on the launching stream.  Chunks are not
    # overlapped across streams here, so each kernel's time is its own (the roofline wants a kernel timed alone).
    NK = L.tlb_kernel_count()
    ms = (C.c_double * NK)()
    cnt = (C.c_uint64 * NK)()
    L.tlb_batch_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    L.tlb_batch_profile(enc._h, 1)
    prof_steps = max(1, min(args.steps, 2))
    serial_s, _ = timed(step_device, prof_steps, 1, sync_each=False)
    L.tlb_batch_kernel_times(enc._h, ms, cnt)
    L.tlb_batch_profile(enc._h, 0)
    if sampler:
        sampler.stop()
    dev_s = allmax(dev_s)
    audio_s_total = world * n_frames * 1152 / FS
    value = audio_s_total * args.steps / dev_s

    # ---- e2e through the host-buffer entry point (pinned memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        h_pcm_p = L.tlb_host_alloc(pcm_bytes)
        h_out_p = L.tlb_host_alloc(out_bytes)
        if not h_pcm_p or not h_out_p:
            raise SystemExit("bench.py: pinned host allocation failed")
        h_pcm = np.ctypeslib.as_array((C.c_int16 * (pcm_bytes // 2)).from_address(h_pcm_p)).reshape(-1, NCH)
        h_out = np.ctypeslib.as_array((C.c_uint8 * out_bytes).from_address(h_out_p))
        torch.from_numpy(h_pcm).copy_(d_pcm)  # same signal, now on the host
        torch.cuda.synchronize()

        def step_host():
            enc.encode(h_pcm, n_frames=n_frames, out=h_out)

        _, wall = timed(step_host, args.steps, args.warmup, sync_each=True)
        wall = allmax(wall)
        e2e = {"value": audio_s_total * args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": pcm_bytes * world,
               "d2h_bytes_per_step": out_bytes * world}
        # the two paths must agree byte for byte
        same = bool(torch.equal(torch.from_numpy(h_out[:lg * 4096]).to(dev), d_out[:lg * 4096]))
        if not same:
            raise SystemExit("bench.py: host-buffer and device-resident outputs differ")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- parity spot check against the oracle (the checker, never the thing measured)
    import oracle
    chk_f0, chk_n = n_frames // 2, 64
    lead = 2  # leading frames: the oracle starts its history at the segment start; frames `lead`.. then see their true
    #           halo (480 samples for the filterbank and psy model 1, 1632 = two blocks back for psy model 2)
    seg = d_pcm[(chk_f0 - lead) * 1152:(chk_f0 + chk_n + 1) * 1152].cpu().numpy()
    ocfg = oracle.configure(FS, MODE, KBPS, PSY)
    want, _ = oracle.encode(ocfg, seg, lead, lead + chk_n)
    got = d_out[chk_f0 * lg:(chk_f0 + chk_n) * lg].cpu().numpy()
    parity_frames_equal = int((got.reshape(chk_n, -1) == want.reshape(chk_n, -1)).all(axis=1).sum())

    # ---- roofline of the dominant kernel
    L.tlb_batch_kernel_name.restype = C.c_char_p
    L.tlb_batch_kernel_name.argtypes = [C.c_void_p, C.c_int]
    names = [L.tlb_batch_kernel_name(enc._h, k).decode() or "unused%d" % k for k in range(NK)]
    per_kernel = {}
    total_ms = sum(ms[k] for k in range(NK)) or 1.0
    for k in range(NK):
        n_l = max(int(cnt[k]), 1)
        per_kernel[names[k]] = {"launches": int(cnt[k]), "avg_ms": ms[k] / n_l, "share": ms[k] / total_ms}
    top = max(range(NK), key=lambda k: ms[k])
    frames_per_launch = n_frames * (prof_steps + 1) / max(int(cnt[top]), 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    kb, kf = KERNEL_ALG.get(names[top], (0, 0))
    dur_s = per_kernel[names[top]]["avg_ms"] * 1e-3
    achieved_gbs = kb * frames_per_launch / dur_s / 1e9
    dfma, dmuladd = C.c_double(), C.c_double()
    L.tlb_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tlb_fp64_peak(local, C.byref(dfma), C.byref(dmuladd))
    for name, rec in per_kernel.items():  # every kernel against both rooflines (algorithmic bytes / flops per launch)
        kb_k, kf_k = KERNEL_ALG.get(name, (0, 0))
        if rec["launches"] and rec["avg_ms"] > 0:
            fpl = n_frames * (prof_steps + 1) / rec["launches"]
            rec["hbm_gbs"] = kb_k * fpl / (rec["avg_ms"] * 1e-3) / 1e9
            rec["hbm_frac"] = rec["hbm_gbs"] / hbm_peak
            rec["fp64_tflops"] = kf_k * fpl / (rec["avg_ms"] * 1e-3) / 1e12
            rec["fp64_frac_of_no_fma_peak"] = rec["fp64_tflops"] / dmuladd.value if dmuladd.value > 0 else None
    roofline = {"kernel": names[top], "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak,
                "traffic": (NCU_DRAM_BYTES_PER_FRAME.get(names[top]) * frames_per_launch
                            if args.config == "B" and names[top] in NCU_DRAM_BYTES_PER_FRAME else None),
                "traffic_source": "ncu --set full, profiles/ncu_r1_summary.md (per frame x frames per launch)",
                "algorithmic_bytes": kb * frames_per_launch, "peak_source": peak_src,
                "fp64": {"achieved_tflops": kf * frames_per_launch / dur_s / 1e12, "peak_dmul_dadd_tflops": dmuladd.value,
                         "peak_dfma_tflops": dfma.value,
                         "frac_of_no_fma_peak": (kf * frames_per_launch / dur_s / 1e12) / dmuladd.value if dmuladd.value > 0 else None},
                "kernels": per_kernel, "serialised_ms_per_step": serial_s / prof_steps * 1e3,
                "frames_per_launch": frames_per_launch}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        times, kind, nf = cpu_reference_run(args.cpu_sample_seconds, cores)
        cpu_baseline = {"value": nf * 1152 / FS * cores / times[0], "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": "%d processes x %d frames (%.0f s) of signal S1, encode loop only" % (cores, nf, nf * 1152 / FS)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD + ", %.3g h synthetic PCM batch per GPU "
                               "(%d frames); inputs (%.2f GB) larger than L2" % (args.hours, n_frames, pcm_bytes / 1e9),
                   "frames_per_gpu": n_frames, "x_realtime_per_gpu": value / world, "host_affinity_rank0": numa},
        "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
        "clocks": sampler.summary() if sampler else None,
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "parity_check": {"frames": chk_n, "byte_identical_to_oracle": parity_frames_equal},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
