#!/usr/bin/env python
"""bench.py -- FLAC encode throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the encode hot path over the whole synthetic stream of the named config
(BASELINE.json configs[1]: 24-bit stereo 96 kHz, 10 min = 57.6 M samples/channel, 14 063 frames of
4096, reference Config.default).  With N > 1 ranks (torchrun) the stream is N x 10 min and every
rank encodes one contiguous frame range on its own GPU ("weak" scaling; no data-path collective --
frames are independent).  `value` is device-resident whole-job MSamples/s (samples x channels);
`e2e` is the same metric through the C-ABI call with pinned HOST buffers (H2D + encode + D2H in the
timed region).  `--impl reference` times the CPU restatement of the reference (oracle port: the
reference is Zig and cannot be built here) on all host threads.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (bit_depth, sample_rate, seconds)
    "c1_16bit_44k1_60s": (16, 44100, 60),
    "c2_24bit_96k_600s": (24, 96000, 600),
    "c3_32bit_192k_600s": (32, 192000, 600),
    # BASELINE config 4: config 1's stream with LPC subframes up to order 12 (the reference has no LPC: the CPU arm is the
    # oracle's statement of the same LPC specification, zigflac_lpc.h)
    "c4_16bit_44k1_60s_lpc12": (16, 44100, 60),
    # BASELINE config 5: the 10-hour stream in eight contiguous frame-range shards; one shard (75 min, 2.6 GB of PCM)
    # per GPU, so `--gpus 8` under torchrun is the whole stream
    "c5_24bit_96k_10h_shard8": (24, 96000, 4500),
    # not a BASELINE config: config 1's format at bandwidth-config length (development aid)
    "x1_16bit_44k1_3600s": (16, 44100, 3600),
}
LPC_ORDER = {"c4_16bit_44k1_60s_lpc12": 12}
DEFAULT_WORKLOAD = "c2_24bit_96k_600s"
BLOCK = 4096
CHANNELS = 2
METRIC = "encode_msamples_per_s"
UNIT = "MSamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-file", action="store_true", help="skip the whole-file (WAV -> FLAC, MD5 included) leg")
    ap.add_argument("--no-decode", action="store_true", help="skip the decoder leg")
    ap.add_argument("--profile", action="store_true", help="device-resident steps only (for runs under ncu)")
    return ap.parse_args()


def shard_of(total_samples, world, rank):
    """Contiguous frame ranges: rank g gets frames [g*ceil(F/G), ...)."""
    frames = (total_samples + BLOCK - 1) // BLOCK
    per = (frames + world - 1) // world
    f0 = min(per * rank, frames)
    f1 = min(f0 + per, frames)
    s0 = f0 * BLOCK
    s1 = min(f1 * BLOCK, total_samples)
    return f0, f1 - f0, s0, max(s1 - s0, 0)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, ~5 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.reasons = set()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


L2_BYTES = 126 * 1000 * 1000


def load_ncu_capture(workload):
    """The committed ncu capture of the dominant kernel ON THIS WORKLOAD (profiles/*_roofline.json: DRAM bytes, executed
    warp instructions and issue-slot utilisation of one launch), latest file first; None when there is none."""
    pdir = os.path.join(ROOT, "profiles")
    found = None
    if os.path.isdir(pdir):
        for name in sorted(os.listdir(pdir)):
            if not name.endswith("_roofline.json"):
                continue
            try:
                d = json.load(open(os.path.join(pdir, name)))
            except Exception:
                continue
            for entry in (d.get("captures") or [d]):
                if entry.get("workload", "c2_24bit_96k_600s") == workload:
                    found = dict(entry, file="profiles/" + name)
    return found


def make_config(workload, world):
    """The `config` object of the JSON line -- the same for both arms (it names the workload, not the measurement)."""
    bits, rate, seconds = WORKLOADS[workload]
    total_samples = rate * seconds * world
    _, nframes, _, nsamples = shard_of(total_samples, world, 0)
    pcm_bytes = nsamples * CHANNELS * bits // 8
    resident = pcm_bytes < 2 * L2_BYTES
    return {"workload": workload, "bit_depth": bits, "sample_rate": rate, "channels": CHANNELS,
            "block_size": BLOCK, "prediction": ("lpc<=%d + fixed" % LPC_ORDER[workload]) if workload in LPC_ORDER else "fixed",
            "stereo_decorrelation": True, "max_rice_order": 8,
            "max_rice_param": 30, "seconds_per_gpu": seconds, "frames_per_gpu": nframes,
            "parallelism": f"frame-range shards x{world}, no collective",
            "l2_resident": resident, "l2_flush": resident,
            "l2": (f"input {pcm_bytes / 1e6:.0f} MB per GPU per step fits the 126 MB L2: a 256 MB buffer is written between "
                   "timed steps (flush), each step timed by its own CUDA events" if resident else
                   f"input {pcm_bytes / 1e6:.0f} MB per GPU per step (+ the FLAC output) is larger than the 126 MB L2: no flush")}


def measured_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, restated (oracle port), all host threads.  Nothing of the
    product is loaded here: the PCM generator comes from oracle/_ref/libzf_synth.so."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    bits, rate, seconds = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    # bounded sample of the workload per step: the first seconds of the stream, about 0.1 s of work on this host
    sample_seconds = min(seconds, 60 if bits == 16 else 30)
    n = (rate * sample_seconds // BLOCK) * BLOCK
    pcm = oracle_lib.synth_pcm(n, rate, bits)
    cfg = oracle_lib.config(CHANNELS, bits, lpc_order=LPC_ORDER.get(args.workload, 0))
    warmup = max(args.warmup, 1)
    for _ in range(warmup):
        oracle_lib.encode_pcm(pcm, n, cfg, rate, threads=cores)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_lib.encode_pcm(pcm, n, cfg, rate, threads=cores)
    dt = (time.perf_counter() - t0) / steps
    value = n * CHANNELS / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32" if bits < 32 else "int64", "data": "synthetic",
        "config": make_config(args.workload, world),
        "realtime_x": round(n / dt / rate, 1),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step encodes the first {sample_seconds} s of the stream ({n} samples/channel), "
                                   f"frames sharded over {cores} host threads"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is Zig (no toolchain in the image): this is its C restatement (oracle port) on ONE host's cores, "
                "whatever --gpus says -- at N > 1 the ratio compares N GPUs with one CPU box",
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import zigflac_b200 as zf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the encode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    bits, rate, seconds = WORKLOADS[args.workload]
    total_samples = rate * seconds * world  # weak scaling: the stream grows with the rank count
    f0, nframes, s0, nsamples = shard_of(total_samples, world, rank)
    ic_bytes = CHANNELS * bits // 8
    pcm_bytes = nsamples * ic_bytes

    # synthetic PCM of this rank's shard -> pinned host memory -> HBM
    h_pcm = torch.empty(pcm_bytes, dtype=torch.uint8, pin_memory=True)
    threads = max(1, (os.cpu_count() or 1) // max(1, min(world, 8)))
    zf.synth_pcm(nsamples, rate, bits, first_sample=s0, threads=threads, out=h_pcm.numpy())
    d_pcm = h_pcm.to(dev, non_blocking=True)

    enc_cfg = zf.Config(CHANNELS, bits, lpc_order=LPC_ORDER.get(args.workload, 0))
    enc = zf.Encoder(enc_cfg, rate, device_id=local_rank, max_frames_per_batch=max(nframes, 1))
    out_cap = enc.max_batch_bytes(nframes)
    d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
    d_sizes = torch.zeros(max(nframes, 1), dtype=torch.int32, device=dev)
    d_total = torch.zeros(1, dtype=torch.int64, device=dev)
    # a non-default stream: its handle is what the C ABI launches on, and the torch events below record on it
    stream = torch.cuda.Stream(dev)
    torch.cuda.synchronize(dev)

    def step():
        enc.encode_device(d_pcm.data_ptr(), nsamples, f0, d_out.data_ptr(), out_cap, d_sizes.data_ptr(),
                          d_total.data_ptr(), stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    enc.kernel_times()  # drop warm-up records
    flac_bytes = int(d_total.item())

    cfg_obj = make_config(args.workload, world)
    flush = cfg_obj["l2_flush"] and not args.profile
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush else None

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    barrier()
    if flush:
        # the whole working set fits the 126 MB L2: write a 256 MB buffer between steps and time every step on its own
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        with torch.cuda.stream(stream):
            for a, b in evs:
                flush_buf.fill_(1)
                a.record(stream)
                step()
                b.record(stream)
        barrier()
        ms_total = sum(a.elapsed_time(b) for a, b in evs)
    else:
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    sampler.stop_flag.set()
    sampler.join()
    ktimes = enc.kernel_times()
    launches_per_step = enc.last_batch_stats()[1]  # full-frame kernel (+ last-frame kernel + append when the stream ends short)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": round(ms_step, 4), "kernel_ms": ktimes}), flush=True)
        enc.close()
        return

    # ---- end to end through the C ABI with pinned host buffers (H2D + encode + D2H every step) ----
    e2e_steps = args.e2e_steps or max(3, min(args.steps, 10))
    enc2 = zf.Encoder(enc_cfg, rate, device_id=local_rank, max_frames_per_batch=2048)
    h_out = torch.empty(out_cap, dtype=torch.uint8, pin_memory=True)
    h_out_np = h_out.numpy()
    h_pcm_np = h_pcm.numpy()
    got, sizes = enc2.encode_pcm(h_pcm_np, nsamples, f0, out=h_out_np)  # warm-up (allocates the staging slots)
    enc2.encode_pcm(h_pcm_np, nsamples, f0, out=h_out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got, sizes = enc2.encode_pcm(h_pcm_np, nsamples, f0, out=h_out_np)
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_out_bytes = int(got.size) + 4 * int(sizes.size)

    # the bare-copy ceiling of that step on this box at this rank count: the same bytes up and down between the same pinned
    # buffers, upload and download on two streams at once, no kernels (one cudaMemcpyAsync per direction)
    up_s, down_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d_stage = torch.empty(max(int(got.size), 1), dtype=torch.uint8, device=dev)
    h_stage = torch.empty(max(int(got.size), 1), dtype=torch.uint8, pin_memory=True)  # `got` lives in h_out: leave it alone

    def bare_copies():
        with torch.cuda.stream(up_s):
            d_pcm.copy_(h_pcm, non_blocking=True)
        with torch.cuda.stream(down_s):
            h_stage.copy_(d_stage, non_blocking=True)
        up_s.synchronize()
        down_s.synchronize()

    bare_copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        bare_copies()
    copy_s = (time.perf_counter() - t0) / e2e_steps
    tc = torch.tensor([copy_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    copy_s = float(tc.item())

    # ---- whole file: WAV bytes -> FLAC bytes through zf_encode_wav_memory (what `flac in.wav out.flac` does after the read:
    #      parse, MD5 of the PCM on a host thread, chunked encode, STREAMINFO back-patch); the serial MD5 bounds it ----
    e2e_file = None
    try:
        if world == 1 and not args.no_e2e_file and pcm_bytes < (1 << 32) - 64:
            import struct
            fmt_chunk = struct.pack("<HHIIHH", 1, CHANNELS, rate, rate * ic_bytes, ic_bytes, bits)
            head = b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt_chunk) + 8 + pcm_bytes) + b"WAVE" + b"fmt " + \
                struct.pack("<I", len(fmt_chunk)) + fmt_chunk + b"data" + struct.pack("<I", pcm_bytes)
            wav = np.empty(len(head) + pcm_bytes, dtype=np.uint8)
            wav[:len(head)] = np.frombuffer(head, dtype=np.uint8)
            wav[len(head):] = h_pcm_np
            if LPC_ORDER.get(args.workload, 0) == 0:  # the driver mirrors the reference CLI: Config.default, no LPC
                rc, flac = zf.wav_to_flac(wav, devices=[local_rank])  # warm-up (page-locks its chunk buffers)
                file_steps = 2
                file_s = 0.0
                for _ in range(file_steps):  # the C call alone: the result stays in the library's buffer (no Python copy)
                    t0 = time.perf_counter()
                    rc, fb = zf.wav_to_flac_view(wav, devices=[local_rank])
                    file_s += (time.perf_counter() - t0) / file_steps
                    if fb is not None:
                        flac = fb.array.tobytes()
                        fb.close()
                md5 = zf.Md5()
                t0 = time.perf_counter()
                md5.update(h_pcm_np)
                digest = md5.final()
                md5_s = time.perf_counter() - t0
                e2e_file = {"value": round(nsamples * CHANNELS / file_s / 1e6, 2), "unit": UNIT, "ms_per_file": round(file_s * 1e3, 2),
                            "md5_only_ms": round(md5_s * 1e3, 2), "flac_bytes": len(flac) if flac else None, "status": rc,
                            "md5_matches_streaminfo": bool(flac and flac[26:42] == digest),
                            "api": "zf_encode_wav_memory (reader | MD5 thread | encoder | writer pipeline over 2048-frame chunks)",
                            "note": "warm library call; bounded by the serial MD5 of the PCM (md5_only_ms), which the reference computes too"}
                # ... and the CLI as a user runs it: a fresh process (CUDA context creation included), files on tmpfs
                cli = os.path.join(ROOT, "zig-flac_b200", "flac")
                shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
                if os.path.exists(cli):
                    import subprocess
                    fin, fout = os.path.join(shm, f"zf_bench_{os.getpid()}.wav"), os.path.join(shm, f"zf_bench_{os.getpid()}.flac")
                    try:
                        wav.tofile(fin)
                        t0 = time.perf_counter()
                        r = subprocess.run([cli, fin, fout], capture_output=True, timeout=300)
                        cli_s = time.perf_counter() - t0
                        same_file = r.returncode == 0 and flac is not None and open(fout, "rb").read() == flac
                        e2e_file["cli"] = {"wall_ms": round(cli_s * 1e3, 1), "exit": r.returncode, "same_bytes_as_library": bool(same_file),
                                           "note": "`flac in.wav out.flac` in a fresh process: process start, CUDA context, file I/O on tmpfs included"}
                    finally:
                        for f in (fin, fout):
                            if os.path.exists(f):
                                os.remove(f)
            del wav

    except Exception as exc:  # the whole-file leg is an extra: never lose the bench line over it
        e2e_file = {"error": repr(exc)}

    # ---- decoder (extension, N = 1): this step's frames as a FLAC stream -> PCM, device-resident (FLAC and PCM in HBM)
    #      and from / to pinned host buffers; the result must be the PCM that went in ----
    decode = None
    try:
        if world == 1 and not args.no_decode and LPC_ORDER.get(args.workload, 0) == 0:
            si = zf.StreamInfo(rate, CHANNELS, bits, nsamples)
            stream_bytes = np.concatenate([np.frombuffer(b"fLaC" + bytes([0x80, 0, 0, 34]) + bytes(si.bytes()), dtype=np.uint8), got])
            h_flac = torch.empty(stream_bytes.size, dtype=torch.uint8, pin_memory=True)
            h_flac.numpy()[:] = stream_bytes
            d_flac = h_flac.to(dev)
            d_dec = torch.zeros(pcm_bytes, dtype=torch.uint8, device=dev)
            h_dec = torch.empty(pcm_bytes, dtype=torch.uint8, pin_memory=True)
            with zf.Decoder(local_rank) as dec:
                for _ in range(2):
                    nb, dinfo = dec.decode_device(d_flac.data_ptr(), d_flac.numel(), d_dec.data_ptr(), d_dec.numel())
                torch.cuda.synchronize(dev)
                dsteps = 5
                t0 = time.perf_counter()
                for _ in range(dsteps):
                    nb, dinfo = dec.decode_device(d_flac.data_ptr(), d_flac.numel(), d_dec.data_ptr(), d_dec.numel())
                torch.cuda.synchronize(dev)
                dev_s = (time.perf_counter() - t0) / dsteps
                dec.decode(h_flac.numpy(), out=h_dec.numpy())
                t0 = time.perf_counter()
                for _ in range(dsteps):
                    dgot, dinfo2 = dec.decode(h_flac.numpy(), out=h_dec.numpy())
                host_s = (time.perf_counter() - t0) / dsteps
            decode = {"value": round(nsamples * CHANNELS / dev_s / 1e6, 2), "unit": UNIT, "ms_per_stream": round(dev_s * 1e3, 3),
                      "algorithmic_gbs": round((int(stream_bytes.size) + pcm_bytes) / dev_s / 1e9, 1),
                      "kernel_ms": round(dinfo["kernel_ms"], 3), "launches": dinfo["launches"], "frames": dinfo["n_frames"],
                      "e2e": {"value": round(nsamples * CHANNELS / host_s / 1e6, 2), "unit": UNIT, "ms_per_stream": round(host_s * 1e3, 3),
                              "h2d_bytes": int(stream_bytes.size), "d2h_bytes": pcm_bytes},
                      "roundtrip_equals_input": bool(nb == pcm_bytes and torch.equal(d_dec, d_pcm) and bool((dgot == h_pcm_np).all())),
                      "api": "zf_decode_flac_device / zf_decode_flac (scan | one thread per frame | CRC-16 | restore + interleave)",
                      "note": "extension (the reference has no decoder); wall clock of the blocking call, the frame table's round trip "
                              "to the host included; bound by the latency of the serial bit parse of a frame (DESIGN.md section 9)"}
            del d_flac, d_dec, h_dec
    except Exception as exc:  # an extra: never lose the bench line over it
        decode = {"error": repr(exc)}

    # the device-resident result and the host-path result are the same bytes
    same = bool((d_out[:flac_bytes].cpu().numpy() == got).all()) and flac_bytes == int(got.size)

    # totals over ranks
    tot = torch.tensor([nsamples, flac_bytes, pcm_bytes, e2e_out_bytes, args.steps * launches_per_step], dtype=torch.int64,
                       device=dev)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    all_samples, all_flac, all_pcm, all_e2e_out, all_launches = [int(v) for v in tot.tolist()]

    if rank == 0:
        value = all_samples * CHANNELS / (ms_step * 1e-3) / 1e6
        e2e_value = all_samples * CHANNELS / e2e_s / 1e6
        peak, peak_src = measured_peak()
        k_ms = statistics.mean(ktimes) if ktimes else ms_step
        full_frames = nsamples // BLOCK
        # algorithmic bytes of one launch of the dominant kernel: packed PCM in + encoded frame bytes out
        sizes_np = d_sizes[:nframes].cpu().numpy()
        kernel_bytes = full_frames * BLOCK * ic_bytes + int(sizes_np[:full_frames].sum())
        achieved = kernel_bytes / (k_ms * 1e-3) / 1e9
        ncu = load_ncu_capture(args.workload)
        # issue-slot view of the same kernel: executed warp instructions of one launch (ncu capture of this workload) against
        # 4 issue slots per SM per clock over the launch's live duration -- the bound the kernel actually runs into
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        clk = (sampler.summary().get("sm_mhz") or 0) * 1e6
        issue = None
        if ncu and ncu.get("inst_executed") and clk:
            inst = float(ncu["inst_executed"])
            issue = {"inst_per_launch": inst, "issue_frac": round(inst / (k_ms * 1e-3 * clk * sm_count * 4), 4),
                     "issue_active_pct_ncu": ncu.get("issue_active_pct"),
                     "note": "warp instructions of one launch (ncu) / (4 slots x SMs x median SM clock x live kernel time)"}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32" if bits < 32 else "int64", "data": "synthetic",
            "config": cfg_obj,
            "realtime_x": round(all_samples / (ms_step * 1e-3) / rate, 1),
            "compression_ratio": round(all_flac / all_pcm, 4),
            "clocks": sampler.summary(),
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": all_pcm,
                    "d2h_bytes_per_step": all_e2e_out, "ms_per_step": round(e2e_s * 1e3, 3), "steps": e2e_steps,
                    "copy_ceiling_ms": round(copy_s * 1e3, 3), "frac_of_copy_ceiling": round(copy_s / e2e_s, 4),
                    "api": "zf_encode_pcm (pinned host PCM in, pinned host FLAC out, 2048-frame batches, pipeline: upload | encode | download on three streams)"},
            "e2e_file": e2e_file,
            "decode": decode,
            "gpu_launches": all_launches,  # all ranks, timed region of the device-resident arm
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": (ncu or {}).get("dram_bytes_per_launch"),
                         "traffic_source": (ncu or {}).get("file"), "issue": issue,
                         "kernel": ("zf::lpc::zf_encode_stereo_lpc_kernel<%d>" % (bits // 8)) if args.workload in LPC_ORDER else ("zf::v3::zf_encode_stereo_v3_kernel<%d>" % (bits // 8)),
                         "kernel_ms": round(k_ms, 4), "algorithmic_bytes_per_launch": kernel_bytes,
                         "peak_source": peak_src,
                         "note": "bound is nominal: the kernel is integer-issue bound (see `issue`), HBM is mostly idle; DESIGN.md section 4"},
            "parity": {"device_path_equals_host_path": same},
        }
        if not args.no_cpu_baseline and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib
            sample_seconds = min(120, seconds)
            n = (rate * sample_seconds // BLOCK) * BLOCK
            cfg = oracle_lib.config(CHANNELS, bits, lpc_order=LPC_ORDER.get(args.workload, 0))
            sample = h_pcm_np[: n * ic_bytes]
            t0 = time.perf_counter()
            ref, ref_sizes = oracle_lib.encode_pcm(sample, n, cfg, rate, threads=1)
            dt1 = time.perf_counter() - t0
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            oracle_lib.encode_pcm(sample, n, cfg, rate, threads=cores)
            dtn = time.perf_counter() - t0
            nb = int(ref_sizes.sum())
            line["cpu_baseline"] = {
                "value": round(n * CHANNELS / dt1 / 1e6, 2), "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"first {sample_seconds} s of the stream ({n} samples/channel), C restatement of zig-flac (the reference is single-threaded)",
                "all_cores": {"value": round(n * CHANNELS / dtn / 1e6, 2), "cores": cores},
            }
            line["parity"]["gpu_equals_oracle_on_sample"] = bool(nb <= got.size and (got[:nb] == ref).all())
            if isinstance(decode, dict) and "error" not in decode:
                # the independent CPU decoder of the test suite on the first 20 s of the same stream, one core (a
                # deliberately plain implementation -- bit-serial CRCs -- so a reported baseline, not a target)
                dk = min(int(ref_sizes.size), rate * 20 // BLOCK)
                dn = dk * BLOCK
                dbytes = int(ref_sizes[:dk].sum())
                dstream = oracle_lib.wrap_frames(ref[:dbytes], CHANNELS, bits, rate, BLOCK, dn)
                t0 = time.perf_counter()
                dres = oracle_lib.decode(dstream)
                ddt = time.perf_counter() - t0
                decode["cpu_baseline"] = {"value": round(dn * CHANNELS / ddt / 1e6, 2), "unit": UNIT, "cores": 1, "kind": "port",
                                          "ok": dres["rc"] == 0,
                                          "sample": f"first {dk} frames, independent CPU decoder of the test suite (plain, bit-serial CRCs)"}
        print(json.dumps(line), flush=True)
    enc.close()
    enc2.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
