#!/usr/bin/env python
"""Headline benchmark: images/sec of CNN-LSTM greedy decode @ 3x64x320 (BASELINE.json
configs[1]: batch 1024 per GPU, max_len 150, bf16, V=512, E=H=256, L=1, random init).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K --warmup W   # CPU oracle port of the reference

A "step" = one pass of the hot path over one batch of synthetic images: encoder +
150-step on-device greedy decode (+ the token all-gather when N > 1).  `value` is
whole-job images/s with the inputs resident in HBM; `e2e` is the same metric through the
public module API with pinned HOST buffers (H2D of the images and D2H of the token ids
inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(model_type="cnn_lstm", vocab_size=512, embedding_dim=256, hidden_dim=256, lstm_layers=1,
           attention=True, img_height=64, img_width=320, channels=3)
START, END, MAX_LEN = 1, 2, 150
METRIC = "images/sec greedy decode @320x64 (CNN-LSTM encoder + 150-step attention-LSTM greedy loop)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def bind_to_gpu_numa(index):
    """Pin this rank's host threads (and thereby the first-touch placement of its pinned staging buffers) to the CPU
    cores NVML reports as local to GPU `index`, when the process is allowed to run there; else leave the affinity
    alone.  Returns a short description for the bench line."""
    info = {"allowed_cpus": len(os.sched_getaffinity(0))}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        ideal = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        both = ideal & allowed
        info["gpu_local_cpus"] = len(ideal)
        if both and both != allowed:
            os.sched_setaffinity(0, both)
            info["bound_to"] = "%d cores local to GPU %d" % (len(both), index)
        else:
            info["bound_to"] = "unchanged (%s)" % ("already local" if both else "GPU-local cores not in the allowed set")
    except Exception as e:                                   # binding is an optimisation, never a failure
        info["bound_to"] = "unchanged (%s)" % type(e).__name__
    return info


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs: ONE long-running
    `nvidia-smi -lms` child (no fork per sample, nothing competing with the launch thread)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            out = ""
        for ln in out.splitlines():
            cols = [c.strip() for c in ln.split(",")]
            if len(cols) >= 6:
                self.rows.append(cols)

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# algorithmic work of each named kernel (per launch), DESIGN.md section "Kernels and rooflines"
# ---------------------------------------------------------------------------------------------
def kernel_work(name, B, steps):
    """(bound, algorithmic work per launch): FLOPs for the tensor-bound convs / FC, bytes for the
    HBM-bound kernels.  conv1 (K = 27) is memory bound: fp32 image in + bf16 pooled map
    out.  Decode: SURVEY 8d's per-step streaming figure W + B*S times the steps run."""
    E = H = 256; V = 512
    conv = {1: (3, 32, 64, 320), 2: (32, 64, 32, 160), 3: (64, 128, 16, 80)}
    if name.startswith("cnn.conv1_u8in"):
        return "hbm", B * (3 * 64 * 320 * 1 + 32 * 160 * 32 * 2)
    if name.startswith("cnn.conv1_bf16in"):
        return "hbm", B * (3 * 64 * 320 * 2 + 32 * 160 * 32 * 2)
    if name.startswith("cnn.conv1_bf16"):
        return "hbm", B * (3 * 64 * 320 * 4 + 32 * 160 * 32 * 2)
    if name.startswith("cnn.conv"):
        i = int(name[len("cnn.conv")])
        ci, co, h, w = conv[i]
        return "tensor", 2.0 * B * h * w * co * ci * 9
    if name.startswith("cnn.fc"):
        return "tensor", 2.0 * B * E * 40960
    if name.startswith("cnn.pool"):
        i = int(name[len("cnn.pool")])
        ci, co, h, w = conv[i]
        return "hbm", B * co * h * w * 4 * 1.25
    if name.startswith("dec.greedy") or name.startswith("dec.sample") or name.startswith("dec.beam"):
        W = 2 * (4 * H * (2 * E + H) + 8 * H + V * H + V)       # bf16 weights
        S = 16 * H + 2 * E + 2 * E + 8                          # per-sequence state traffic
        return "hbm", float(steps) * (W + B * S)
    return None, 0.0


def run_ours(args):
    import torch
    import torch.distributed as dist
    import i2l_import
    pkg = i2l_import.load()
    N = pkg._native
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    binding = bind_to_gpu_numa(local)                    # before any pinned allocation
    if world > 1:
        # NCCL may print its version banner to stdout at communicator creation (NCCL_DEBUG=VERSION on
        # some boxes): keep stdout to the one JSON line by pointing fd 1 at stderr until the first collective ran
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from hmer_img2latex_b200.dist import TokenExchange, gather_tokens

    torch.manual_seed(0)
    model = pkg.Seq2SeqModel("cnn_lstm", CFG["vocab_size"],
                             dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                             dict(hidden_dim=256, lstm_layers=1, attention=True), precision=args.precision)
    model = model.to(dev).eval()
    B = args.batch
    g = torch.Generator().manual_seed(100 + rank)
    # NB different batches, used in turn: NB x B x 3x64x320 exceeds the 126 MB L2 in every dtype.
    # uint8 (default): raw RGB pixels = clamp(randn)/2 mapped to 0..255, normalised on the device as
    # x/255*2-1 (Predictor._prepare_image, training/predictor.py:441-446) inside conv1; bf16 / fp32: the
    # already normalised tensors (the reference model's own input dtype is fp32).
    NB = 3
    x_master = [torch.randn(B, 3, 64, 320, generator=g) for _ in range(NB)]

    def to_dtype(xm, nm):
        if nm == "uint8":
            return ((xm.clamp(-2, 2) / 2 + 1) * 127.5).round().to(torch.uint8)
        return xm.to({"bf16": torch.bfloat16, "fp32": torch.float32}[nm])

    x_host = [to_dtype(xm, args.input_dtype).pin_memory() for xm in x_master]
    x_dev = [xh.to(dev) for xh in x_host]
    lib = N.lib()

    # N > 1: the token all-gather runs as direct peer stores into symmetric memory (dist.TokenExchange,
    # csrc/token_exchange.cu), pipelined one step behind the compute; NCCL all_gather_into_tensor (gather_tokens) is the
    # checked reference of that path and the fallback when symmetric memory cannot be set up on this box.
    xchg, exchange_kind = None, "none"
    if world > 1 and not args.nccl_gather:
        try:
            xchg = TokenExchange(B * world, MAX_LEN + 1, dev)
            exchange_kind = ("p2p stores into symmetric memory + flags (i2l_token_exchange_*) on a side stream, released beside "
                             "the next batch's decode kernel (which leaves 20 SMs idle), results one step behind")
        except Exception as e:
            print("TokenExchange unavailable (%r): NCCL all-gather instead" % (e,), file=sys.stderr)
    if world > 1 and xchg is None:
        exchange_kind = "NCCL all_gather_into_tensor on the compute stream"

    def step(inp, xc=None, rows=None):
        """one pass of the hot path over one batch; N > 1: the shard result is staged for the p2p exchange, which is
        released (TokenExchange.kick) between the NEXT batch's encoder and decode and runs beside that decode"""
        xc = xchg if xc is None else xc
        if rows is not None:
            inp = inp[:rows]
        enc = model.encoder.forward_u8(inp) if inp.dtype == torch.uint8 else model.encoder(inp)
        side = args.xchg_mode == "side"
        prev = xc.kick() if (world > 1 and xc is not None and side) else None
        tokens, lengths, steps = model.decoder.greedy(enc, START, END, MAX_LEN, 1.0, N.STOP_ALL_END_SAME_STEP)
        if world > 1 and args.xchg_mode != "none":
            if xc is not None and side:
                xc.stage(tokens, lengths, steps)
                return prev if prev is not None else (tokens, lengths, steps)
            if xc is not None:
                got = xc.read() if xc.read_seq != xc.seq else None
                xc.write(tokens, lengths, steps)
                return got if got is not None else (tokens, lengths, steps)
            tokens, lengths, steps = gather_tokens(tokens, lengths, steps, inp.shape[0] * world)
        return tokens, lengths, steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def e2e_run(hosts, n):
        """n batches through Seq2SeqModel.greedy_stream from pinned HOST buffers; returns seconds."""
        def host_batches(k):
            for i in range(k):
                yield hosts[i % len(hosts)]
        rb = "global" if rank == 0 else "shard"      # rank 0 hands the job's (global) result to the host; the others their rows
        for _ in model.greedy_stream(host_batches(6 if world == 1 else 12), START, END, MAX_LEN, exchange=xchg, exchange_readback=rb):
            pass
        barrier()
        t0 = time.perf_counter()
        last = None
        for last in model.greedy_stream(host_batches(n), START, END, MAX_LEN, exchange=xchg, exchange_readback=rb):
            pass   # N > 1: exchange on the device + D2H (rank 0: the GLOBAL id matrix) inside the timed region
        barrier()
        return time.perf_counter() - t0, last

    def h2d_only(hosts, n):
        """ceiling of the end-to-end path: the same pinned batches copied to the device, nothing else; seconds"""
        dst = [torch.empty_like(hosts[0], device=dev) for _ in range(2)]
        cs = torch.cuda.Stream(dev)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(cs):
            for i in range(n):
                dst[i & 1].copy_(hosts[i % len(hosts)], non_blocking=True)
        cs.synchronize()
        barrier()
        return time.perf_counter() - t0

    with torch.no_grad():
        # clocks / throttle reasons are sampled by one `nvidia-smi -lms 50` child from the warm-up on (the tool needs
        # ~0.1-0.3 s before its first line, longer when 8 ranks start one each) through both timed regions
        sampler = ClockSampler(local); sampler.start()
        # N > 1: the exchange pipeline (side stream, result ring, allocator pool) reaches its steady state after a few
        # dozen steps; they are run untimed on top of the requested warm-up
        for i in range(max(args.warmup, 3) + (32 if world > 1 else 0)):
            step(x_dev[i % NB])
        if xchg is not None:
            # parity of the p2p exchange with the library collective on the same shard results (every rank checks)
            xchg.flush()
            enc = model.encoder.forward_u8(x_dev[0]) if x_dev[0].dtype == torch.uint8 else model.encoder(x_dev[0])
            t_, l_, s_ = model.decoder.greedy(enc, START, END, MAX_LEN, 1.0, N.STOP_ALL_END_SAME_STEP)
            ref_t, ref_l, ref_s = gather_tokens(t_, l_, s_, B * world)
            xchg.write(t_, l_, s_)
            got_t, got_l, got_s = xchg.read()
            xchg.check()
            assert torch.equal(got_t, ref_t) and torch.equal(got_l, ref_l) and int(got_s) == int(ref_s), \
                "p2p token exchange differs from the NCCL all-gather"
        # everything that costs host time (timer reset: event destruction / creation) happens BEFORE the barrier: at N > 1
        # every rank's region ends when the slowest rank's last shard has arrived, so a rank that leaves the start line
        # late charges its delay to all of them
        lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if world > 1:
            # device-side start line: the ranks leave the host barrier up to a millisecond apart, and with the exchange every
            # rank's region ends when the LAST rank's shard has arrived -- so the start events are recorded behind one
            # more collective on the compute stream, whose kernels finish together on all GPUs
            dist.all_reduce(torch.zeros(1, device=dev))
        # ---- device-resident timed region ------------------------------------------------
        l0 = lib.i2l_launch_count()
        t_host = [time.perf_counter() * 1e3] * 4
        e0.record()
        t_host[1] = time.perf_counter() * 1e3
        for i in range(args.steps):
            out = step(x_dev[i % NB])
        t_host[2] = time.perf_counter() * 1e3
        if xchg is not None:
            fl = xchg.flush()                                 # the last steps' global results: inside the timed region
            out = fl[-1] if fl else out
        e1.record()
        barrier()
        t_host[3] = time.perf_counter() * 1e3
        launches = lib.i2l_launch_count() - l0
        lib.i2l_prof_enable(0)
        ms = e0.elapsed_time(e1)
        prof = N.prof_results()
        steps_run = int(out[2].item())
        # ---- end-to-end through the public API with host buffers --------------------------
        # Seq2SeqModel.greedy_stream: every step copies its own images from pinned host memory
        # (the copy of step i+1 overlaps the compute of step i) and reads the token ids back.
        # (a 20-batch stream lasts ~25 ms: one host hiccup of a few ms moves it by 10 %; the end-to-end figure is taken over
        # at least 60 batches and reported with its own count)
        e2e_steps = max(args.steps, 60)
        e2e_s, (tok_h, lens_h, _) = e2e_run(x_host, e2e_steps)
        h2d_s = h2d_only(x_host, e2e_steps)
        # ---- strong scaling (N > 1): BASELINE configs[1]'s global batch of 1024 split over the ranks ----------------
        strong = None
        if world > 1 and B % world == 0 and not args.no_extras:
            Bs = B // world
            xs = None
            try:
                xs = TokenExchange(B, MAX_LEN + 1, dev) if xchg is not None else None
            except Exception:
                xs = None
            for i in range(3):
                step(x_dev[i % NB], xs, Bs)
            if xs is not None:
                xs.flush()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(args.steps):
                step(x_dev[i % NB], xs, Bs)
            if xs is not None:
                xs.flush()
            s1.record()
            barrier()
            strong = s0.elapsed_time(s1)
        # short runs: keep the same load up until the sampler has lines.  The count depends on the arguments only --
        # every rank must run the same number of steps (a step ends in the token all-gather)
        for i in range(max(0, 600 - 2 * args.steps)):
            step(x_dev[i % NB])
        if xchg is not None:
            xchg.flush(); xchg.check()
        torch.cuda.synchronize()
        sampler.stop()
        # the same call with the other host element types (reported next to the headline e2e)
        e2e_other = {}
        if world == 1 and not args.no_extras:
            for nm in ("uint8", "bf16", "fp32"):
                if nm == args.input_dtype:
                    continue
                hosts = [to_dtype(xm, nm).pin_memory() for xm in x_master[:2]]
                sec, _ = e2e_run(hosts, args.steps)
                e2e_other[nm] = {"value": round(B * args.steps / sec, 1), "unit": "images/s",
                                 "h2d_bytes_per_step": hosts[0].numel() * hosts[0].element_size()}
                del hosts
        # ---- BASELINE configs[2]: beam search, beam 5, batch 512 per GPU --------------------
        beam = None
        if not args.no_extras:
            Bb, K = args.beam_batch, 5
            enc_of = lambda inp: model.encoder.forward_u8(inp) if inp.dtype == torch.uint8 else model.encoder(inp)
            encb = enc_of(x_dev[0][:Bb])
            for _ in range(2):
                model.decoder.beam(encb, START, END, MAX_LEN, K)
            barrier()
            lib.i2l_prof_reset(); lib.i2l_prof_enable(1)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            nb = max(2, min(args.steps, 5))
            b0.record()
            for i in range(nb):
                encb = enc_of(x_dev[i % NB][:Bb])
                model.decoder.beam(encb, START, END, MAX_LEN, K)
            b1.record()
            barrier()
            lib.i2l_prof_enable(0)
            bms = b0.elapsed_time(b1) / nb
            bprof = N.prof_results()
            bdec = sum(v[1] for k, v in bprof.items() if k.startswith("dec.")) / nb
            beam = [bms, bdec, Bb, K, {k: round(v[1] / nb, 4) for k, v in sorted(bprof.items())}]
        resnet = None
        nxt = None
        if not args.no_extras:
            try:
                resnet = resnet_lines(pkg, dev, B, world=world)
            except Exception as e:                       # an extra line must never take the headline down
                resnet = {"error": repr(e)[:300]}
            try:
                nxt = next_row_lines(pkg, dev, B) if world == 1 else None
            except Exception as e:
                nxt = {"error": repr(e)[:300]}
    if world > 1 and os.environ.get("I2L_BENCH_RANK_DEBUG"):
        print("rank %d: timed region %.3f ms (host: start %.3f, enqueued %.3f, done %.3f ms after rank-local barrier exit), e2e %.3f ms, kernels %s"
              % (rank, ms, t_host[1] - t_host[0], t_host[2] - t_host[0], t_host[3] - t_host[0], e2e_s * 1e3,
                 {k: (v[0], round(v[1] / max(v[0], 1), 4)) for k, v in sorted(prof.items())}),
              file=sys.stderr, flush=True)
    tms = torch.tensor([ms, e2e_s * 1e3, beam[0] if beam else 0.0, h2d_s * 1e3, strong or 0.0], device=dev,
                       dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, beam_ms, h2d_ms, strong_ms = (float(v) for v in tms)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B * args.steps / (ms / 1e3)
    pk = peaks()
    # dominant kernel by measured time
    roof = None
    if prof:
        name, (cnt, tot) = max(prof.items(), key=lambda kv: kv[1][1])
        bound, work = kernel_work(name, B, steps_run)
        if bound:
            per_launch_s = tot / cnt / 1e3
            if bound == "hbm":
                ach, peak, unit = work / per_launch_s / 1e9, pk["hbm"], "GB/s"
            else:
                ach, peak, unit = work / per_launch_s / 1e12, pk["tf_sust"], "TFLOP/s"
            roof = {"kernel": name, "bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit,
                    "frac": round(ach / peak, 4), "traffic": None, "peak_source": pk["src"],
                    "launch_ms": round(tot / cnt, 4), "share_of_step": round(tot / ms, 4)}
    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)
    try:
        tr_file = "r2_traffic.json" if os.path.exists(os.path.join(ROOT, "profiles", "r2_traffic.json")) else "r1_traffic.json"
        tr = json.load(open(os.path.join(ROOT, "profiles", tr_file)))["kernels"]
        if roof and roof["kernel"] in tr and B == 1024:
            roof["traffic"] = tr[roof["kernel"]]["dram_bytes_per_launch"]
            roof["traffic_unit"] = "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, profiles/%s)" % tr_file
    except Exception:
        pass
    rooflines = {}
    for kname, (cnt, tot) in sorted(prof.items()):
        bound, work = kernel_work(kname, B, steps_run)
        if bound:
            t = tot / cnt / 1e3
            ach = work / t / (1e9 if bound == "hbm" else 1e12)
            peak = pk["hbm"] if bound == "hbm" else pk["tf_sust"]
            rooflines[kname] = {"bound": bound, "achieved": round(ach, 1), "frac": round(ach / peak, 3),
                                "ms": round(tot / cnt, 4)}
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: CNN-LSTM greedy decode, batch %d per GPU, 3x64x320 images, "
                               "max_len 150, V=512, E=H=256, L=1, random init" % B,
                   "global_batch": B * world, "parallelism": "dp%d (batch-sharded, token all-gather)" % world,
                   "token_exchange": exchange_kind, "host_binding": binding,
                   "untimed_steps_before_the_timed_region": max(args.warmup, 3) + (32 if world > 1 else 0),
                   "l2_policy": "%d input batches used in turn (%d x %.0f MB of images) exceed the 126 MB L2; "
                                "the bf16 activations written per step (587 MB) flush it as well"
                                % (NB, NB, x_dev[0].numel() * x_dev[0].element_size() / 1e6),
                   "input": {"uint8": "raw uint8 RGB pixels, x/255*2-1 (Predictor._prepare_image) fused into conv1",
                             "bf16": "normalised bf16 tensors", "fp32": "normalised fp32 tensors"}[args.input_dtype],
                   "decode_steps_run": steps_run, "host_input_dtype": args.input_dtype},
        "us_per_decode_step": None,
        "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items())},
        "roofline": roof,
        "roofline_all_kernels": rooflines,
        "e2e": {"value": round(world * B * e2e_steps / (e2e_ms / 1e3), 1), "unit": "images/s", "steps": e2e_steps,
                "h2d_bytes_per_step": x_host[0].numel() * x_host[0].element_size(),
                "d2h_bytes_per_step": tok_h.numel() * 8 + lens_h.numel() * 4, "host_dtype": args.input_dtype,
                "h2d_only_ceiling": {"value": round(world * B * e2e_steps / (h2d_ms / 1e3), 1), "unit": "images/s",
                                     "GB_per_s_per_gpu": round(x_host[0].numel() * x_host[0].element_size() * e2e_steps
                                                               / (h2d_ms / 1e3) / 1e9, 2),
                                     "note": "the same pinned host batches copied to the device and nothing else "
                                             "(max over ranks): the PCIe / host-memory bound of e2e"},
                "frac_of_h2d_ceiling": round(h2d_ms / e2e_ms, 3)},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if strong_ms > 0:
        line["strong_scaling"] = {
            "workload": "BASELINE configs[1] with the GLOBAL batch fixed at %d (%d images per GPU), token exchange "
                        "included" % (B, B // world),
            "value": round(B * args.steps / (strong_ms / 1e3), 1), "unit": "images/s", "scaling": "strong",
            "ms_per_step": round(strong_ms / args.steps, 4),
            "note": "the 150-step decode is a latency chain whose step time does not shrink with the batch (one "
                    "cluster of 4 SMs per 32 sequences either way); only the encoder scales with 1/N"}
    if e2e_other:
        line["e2e_other_host_dtypes"] = e2e_other
    if beam:
        _, bdec, Bb, K, bk = beam
        W = 2 * (4 * 256 * 768 + 8 * 256 + 512 * 256 + 512)
        S = 16 * 256 + 2 * 256 + 2 * 256 + 8 + 16 * 256      # + the K-way h/c gather
        bbytes = MAX_LEN * (W + Bb * K * S)
        line["beam5"] = {"value": round(world * Bb / (beam_ms / 1e3), 1), "unit": "images/s", "batch_per_gpu": Bb,
                         "beam": K, "ms_per_step": round(beam_ms, 3),
                         "us_per_decode_step": round(bdec / MAX_LEN * 1e3, 3), "kernels_ms_per_step": bk,
                         "roofline": {"bound": "hbm", "achieved": round(bbytes / (bdec / 1e3) / 1e9, 1),
                                      "peak": pk["hbm"], "unit": "GB/s",
                                      "frac": round(bbytes / (bdec / 1e3) / 1e9 / pk["hbm"], 4)},
                         "workload": "BASELINE configs[2]: CNN-LSTM beam search, beam 5, batch %d per GPU, max_len 150" % Bb}
    if resnet:
        line.update(resnet)
    if nxt:
        line["next_rows"] = nxt
    dk = [v[1] for k, v in prof.items() if k.startswith("dec.")]
    if dk and steps_run:
        line["us_per_decode_step"] = round(sum(dk) / args.steps / steps_run * 1e3, 3)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(model, sample_batch=32, budget_s=20.0)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def resnet_lines(pkg, dev, B, reps=3, world=1):
    """BASELINE configs[3] / configs[4] (extra lines next to the headline; N > 1: B images per GPU, ids all-gathered over
    NCCL inside the timed step, max time over ranks, whole-job images/s):
    ResNet18-LSTM greedy decode over width-bucketed images (widths uniform in {128,160,...,800}),
    ResNet50-LSTM temperature / top-k / top-p sampling at 3x64x320.  E = H = 256, L = 1, V = 512,
    bf16 tcgen05 trunk (resnet_bf16.cu); images are created on the device (no e2e figure here)."""
    import torch
    N = pkg._native
    lib = N.lib()
    out = {}
    mk = lambda name: pkg.Seq2SeqModel("resnet_lstm", CFG["vocab_size"],
                                       dict(img_height=64, img_width=320, channels=3, model_name=name, embedding_dim=256),
                                       dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").to(dev).eval()
    g = torch.Generator().manual_seed(5)
    # ---- configs[3]: width buckets
    m = mk("resnet18")
    widths = (torch.randint(0, 22, (B,), generator=g) * 32 + 128).tolist()
    buckets = {}
    for w in widths:
        buckets[w] = buckets.get(w, 0) + 1
    xs = {w: torch.randn(n, 3, 64, w, device=dev) for w, n in sorted(buckets.items())}

    import torch.distributed as dist
    from hmer_img2latex_b200.dist import gather_tokens

    def gathered(res):
        return gather_tokens(res[0], res[1], res[2], B * world) if world > 1 else res

    def step18():
        enc = torch.cat(m.encoder.forward_buckets(list(xs.values()), n_streams=int(os.environ.get('I2L_BUCKET_STREAMS', '4'))), 0)
        return gathered(m.decoder.greedy(enc, START, END, MAX_LEN, 1.0, N.STOP_ALL_END_SAME_STEP))

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:                                      # max over ranks
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        lib.i2l_prof_reset(); lib.i2l_prof_enable(1)       # separate pass for the per-kernel split (event records cost host time)
        fn()
        torch.cuda.synchronize()
        lib.i2l_prof_enable(0)
        prof = N.prof_results()
        return ms, {k: round(v[1], 4) for k, v in sorted(prof.items())}

    ms, prof = timed(step18)
    fl = 4.627e6 * sum(widths)
    out["resnet18_bucketed_greedy"] = {
        "workload": "BASELINE configs[3]: ResNet18-LSTM greedy decode, %d images, widths uniform in {128..800 step 32} "
                    "(%d buckets), max_len 150" % (B, len(buckets)),
        "value": round(world * B / ms * 1e3, 1), "unit": "images/s", "n_gpus": world, "ms_per_step": round(ms, 3),
        "decode_ms": round(sum(v for k, v in prof.items() if k.startswith("dec.")), 3),
        "note": "buckets replayed as CUDA graphs on 4 streams (ResNetEncoder.forward_buckets)"}
    r = out["resnet18_bucketed_greedy"]
    r["encoder_ms"] = round(ms - r["decode_ms"], 3)
    r["encoder_tflops"] = round(fl / r["encoder_ms"] / 1e9, 1)
    r["encoder_frac_of_bf16_peak"] = round(r["encoder_tflops"] / peaks()["tf_sust"], 3)
    del m, xs
    # ---- configs[4]: ResNet50 + sampling
    m = mk("resnet50")
    x = torch.randn(B, 3, 64, 320, device=dev)

    def step50():
        enc = m.encoder(x)
        return gathered(m.decoder.sample(enc, START, END, MAX_LEN, temperature=0.8, top_k=50, top_p=0.9, seed=1))

    ms, prof = timed(step50)
    enc_ms = sum(v for k, v in prof.items() if k.startswith("rn."))
    fl = 10.43e6 * 320 * B
    out["resnet50_sampling"] = {
        "workload": "BASELINE configs[4]: ResNet50-LSTM sampling (temperature 0.8, top_k 50, top_p 0.9), "
                    "batch %d per GPU on %d GPU(s), 3x64x320, max_len 150" % (B, world),
        "value": round(world * B / ms * 1e3, 1), "unit": "images/s", "n_gpus": world, "ms_per_step": round(ms, 3),
        "encoder_ms": round(enc_ms, 3), "encoder_tflops": round(fl / enc_ms / 1e9, 1),
        "encoder_frac_of_bf16_peak": round(fl / enc_ms / 1e9 / peaks()["tf_sust"], 3),
        "decode_ms": round(ms - enc_ms, 3)}
    del m, x
    # ---- the reference's SHIPPED configuration (img2latex/configs/config.yaml:45-50: embedding_dim 512, hidden_dim 512,
    # lstm_layers 2; encoder.py:50-64: grey 1x64x800 images) -> two extra lines:
    #   (a) CNN encoder at 1x64x800 (tcgen05 path, cnn_bf16.cu) + the headline decoder: the serving shape of Predictor
    #   (b) the 512 / 512 / 2 decoder alone (graph-replayed stream-ordered loop with the fused gate-GEMM + cell launch)
    m = pkg.Seq2SeqModel("cnn_lstm", CFG["vocab_size"], dict(img_height=64, img_width=800, channels=1, embedding_dim=256),
                         dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").to(dev).eval()
    px = torch.randint(0, 256, (B, 1, 64, 800), dtype=torch.uint8, device=dev)

    def step800():
        return gathered(m.decoder.greedy(m.encoder.forward_u8(px), START, END, MAX_LEN, 1.0, N.STOP_ALL_END_SAME_STEP))

    ms, prof = timed(step800)
    enc_ms = sum(v for k, v in prof.items() if k.startswith("cnn."))
    fl = B * 2.0 * (32 * 9 * 1 * 64 * 800 + 64 * 9 * 32 * 32 * 400 + 128 * 9 * 64 * 16 * 200 + 256 * 128 * 8 * 100)
    out["cnn_1x64x800_greedy"] = {
        "workload": "reference default / serving shape (encoder.py:50-64, predictor.py:409-414): CNN-LSTM greedy decode, "
                    "%d grey 1x64x800 uint8 images per GPU on %d GPU(s), max_len 150" % (B, world),
        "value": round(world * B / ms * 1e3, 1), "unit": "images/s", "n_gpus": world, "ms_per_step": round(ms, 3),
        "encoder_ms": round(enc_ms, 3), "encoder_tflops": round(fl / enc_ms / 1e9, 1),
        "encoder_frac_of_bf16_peak": round(fl / enc_ms / 1e9 / peaks()["tf_sust"], 3),
        "kernels_ms": {k: v for k, v in prof.items() if k.startswith("cnn.") or k.startswith("dec.")}}
    del m, px
    dec = pkg.LSTMDecoder(CFG["vocab_size"], 512, 512, MAX_LEN, 2, 0.0, True, precision="bf16").to(dev).eval()
    enc = torch.relu(torch.randn(B, 512, device=dev))
    ms, prof = timed(lambda: gathered(dec.greedy(enc, START, END, MAX_LEN, 1.0, N.STOP_NONE)))
    dec.streamed = True                                         # I2L_BF16_STREAMED: the stream-ordered loop, same weights
    ms_streamed, _ = timed(lambda: gathered(dec.greedy(enc, START, END, MAX_LEN, 1.0, N.STOP_NONE)))
    dec.streamed = False
    fl_step = B * 11.01e6                                       # live FLOPs per sequence and step @512/512/2
    out["decoder_512_512_2_greedy"] = {
        "workload": "the reference's shipped decoder (configs/config.yaml:45-50: E = H = 512, 2 LSTM layers), greedy, "
                    "%d sequences per GPU on %d GPU(s), 150 steps, V = 512" % (B, world),
        "value": round(world * B / ms * 1e3, 1), "unit": "sequences/s", "n_gpus": world, "ms_per_step": round(ms, 3),
        "us_per_decode_step": round(ms / MAX_LEN * 1e3, 2),
        "roofline": {"bound": "tensor", "achieved": round(fl_step / (ms / MAX_LEN * 1e-3) / 1e12, 1),
                     "peak": peaks()["tf_sust"], "unit": "TFLOP/s",
                     "frac": round(fl_step / (ms / MAX_LEN * 1e-3) / 1e12 / peaks()["tf_sust"], 3)},
        "kernels_ms": {k: v for k, v in prof.items() if k.startswith("dec.")},
        "stream_ordered_us_per_decode_step": round(ms_streamed / MAX_LEN * 1e3, 2),
        "note": "decode_wide.cu: ONE cooperative kernel runs all 150 steps on the whole GPU (128-sequence x 128-gate-column "
                "tiles with static ownership, one counter per 128-sequence block instead of launches, weights streamed "
                "from L2 -- 6.5 MB of bf16 weights do not fit a cluster); stream_ordered_* = the CUDA-graph loop of "
                "L + 2 launches per step it replaces (I2L_BF16_STREAMED)"}
    return out


def next_row_lines(pkg, dev, B):
    """SURVEY 8f rows next to the headline (one GPU): device-side image preparation (f1), evaluation metrics
    (f3) and the teacher-forced decoder pass (f4), each with its algorithmic-bytes roofline and the reference's
    CPU arithmetic (Pillow / the pure-Python port) timed on a bounded sample of the same work."""
    import ctypes as C
    import time
    import numpy as np
    import torch
    N = pkg._native
    lib = N.lib()
    pk = peaks()
    out = {}

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- f1: ragged grey images -> 64 x 800 (the size load_image / Predictor use), LANCZOS + pad / crop
    P = pkg.preprocess
    rng = np.random.default_rng(0)
    imgs = []
    for _ in range(B):
        h = int(rng.integers(32, 128)); w = int(h * rng.uniform(2.0, 14.0))
        imgs.append(rng.integers(0, 256, size=(h, w), dtype=np.uint8))
    plan = P.ResizePlan(imgs, 64, 800)
    dplan = plan.host.to(dev)
    o = torch.empty(B, 1, 64, 800, dtype=torch.uint8, device=dev)
    ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device=dev)

    def resize():
        N.check(lib.i2l_resize_pad_u8(N.ptr(dplan), C.c_void_p(plan.plan_ptr), C.c_void_p(dplan.data_ptr() + plan.pixel_bytes),
                                      N.ptr(o), N.ptr(ws), ws.numel(), N.stream_ptr(dev)), "i2l_resize_pad_u8")
    ms = timed(resize)
    src = sum(a.size for a in imgs)
    alg = src + 2 * (plan.workspace_bytes - 512) + o.numel()          # source + intermediate write/read + output
    for _ in range(2):                                   # both pinned staging slots exist before the clock starts
        P.ResizePlan(imgs, 64, 800).run(dev)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(5):
        P.ResizePlan(imgs, 64, 800).run(dev)
    torch.cuda.synchronize(dev)
    e2e = (time.perf_counter() - t0) / 5
    row = {"workload": "SURVEY 8f-1: %d ragged grey images (h 32-127, aspect 2-14) -> 64x800, Pillow LANCZOS + pad/crop, "
                       "bit-exact" % B,
           "value": round(B / ms * 1e3, 1), "unit": "images/s", "ms": round(ms, 4),
           "roofline": {"bound": "hbm", "achieved": round(alg / ms / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                        "frac": round(alg / ms / 1e6 / pk["hbm"], 4), "algorithmic_bytes": int(alg)},
           "e2e": {"value": round(B / e2e, 1), "unit": "images/s", "h2d_bytes_per_step": int(plan.host.numel()),
                   "note": "host arrays -> filter-weight plan on the host cores -> one H2D -> two launches"}}
    try:
        from PIL import Image
        n = min(B, 128)
        t0 = time.perf_counter()
        for a in imgs[:n]:
            im = Image.fromarray(a, "L")
            nw = int(round(64 * (a.shape[1] / a.shape[0])))
            r = im.resize((nw, 64), Image.Resampling.LANCZOS)
            if nw < 800:
                q = Image.new("L", (800, 64), 255); q.paste(r, (0, 0)); r = q
            elif nw > 800:
                l = (nw - 800) // 2; r = r.crop((l, 0, l + 800, 64))
        row["cpu_baseline"] = {"value": round(n / (time.perf_counter() - t0), 1), "unit": "images/s", "cores": 1,
                               "kind": "reference", "sample": "Pillow (the reference's own resampler) on the first %d images" % n}
    except ImportError:
        pass
    out["preprocess_resize"] = row
    del dplan, o, ws

    # ---- f3: BLEU-4 + Levenshtein counts of B (prediction, target) pairs of ~150 tokens
    M = pkg.metrics
    g = torch.Generator().manual_seed(0)
    T = MAX_LEN + 1
    t = torch.randint(0, 100, (B, T), generator=g)
    p = torch.where(torch.rand(B, T, generator=g) < 0.15, torch.randint(0, 100, (B, T), generator=g), t)
    lt = torch.randint(100, T + 1, (B,), generator=g, dtype=torch.int32)
    lp = (lt - torch.randint(0, 10, (B,), generator=g, dtype=torch.int32)).clamp(min=1)
    pc, tc, lpc, ltc = p.to(dev), t.to(dev), lp.to(dev), lt.to(dev)
    ms = timed(lambda: M.sequence_counts(pc, lpc, tc, ltc))
    cells = float((lp.double() * lt.double()).sum())
    row = {"workload": "SURVEY 8f-3: edit distance + clipped 1..4-gram matches of %d id-sequence pairs (100-151 tokens)" % B,
           "value": round(B / ms * 1e3, 1), "unit": "pairs/s", "ms": round(ms, 4),
           "dp_cells_per_s": round(cells / ms * 1e3, 1),
           "note": "integer DP held in registers / shared memory: neither HBM- nor tensor-bound (16 B of ids per token "
                   "in, 32 B per pair out); the rate is set by dependent shared-memory reads"}
    from oracle import metrics as OM                       # cpu_baseline leg: the reference's pure-Python loops restated
    n = min(B, 16)
    t0 = time.perf_counter()
    OM.calculate_metrics([p[i, : lp[i]].tolist() for i in range(n)], [t[i, : lt[i]].tolist() for i in range(n)])
    row["cpu_baseline"] = {"value": round(n / (time.perf_counter() - t0), 1), "unit": "pairs/s", "cores": 1, "kind": "port",
                           "sample": "first %d pairs, pure-Python port of training/metrics.py" % n}
    out["evaluate_metrics"] = row

    # ---- f4: teacher-forced decoder pass, headline decoder, B x 150 tokens
    m = pkg.Seq2SeqModel("cnn_lstm", CFG["vocab_size"], dict(img_height=64, img_width=320, channels=3, embedding_dim=256),
                         dict(hidden_dim=256, lstm_layers=1, attention=True), precision="bf16").to(dev).eval()
    enc = torch.relu(torch.randn(B, 256, generator=g)).to(dev)
    tg = torch.randint(0, CFG["vocab_size"], (B, MAX_LEN), generator=g).to(dev)
    ms = timed(lambda: m.decoder(enc, tg), reps=10)
    W = 2 * (4 * 256 * 768 + 8 * 256 + 512 * 256 + 512)
    S = 16 * 256 + 2 * 256 + 2 * 256 + 8
    alg = MAX_LEN * (W + B * S) + B * MAX_LEN * 512 * 4                 # SURVEY 8d per-step bytes + the fp32 logits written
    out["teacher_forced_forward"] = {
        "workload": "SURVEY 8f-4: LSTMDecoder.forward (eval), %d x %d tokens, E=H=256, L=1, V=512, bf16 GEMMs" % (B, MAX_LEN),
        "value": round(B * MAX_LEN / ms * 1e3, 1), "unit": "tokens/s", "ms": round(ms, 3),
        "us_per_step": round(ms / MAX_LEN * 1e3, 2),
        "roofline": {"bound": "hbm", "achieved": round(alg / ms / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                     "frac": round(alg / ms / 1e6 / pk["hbm"], 4)},
        "note": "all steps inside the persistent cluster kernel (mode 2: given tokens, logits written per step, no "
                "token exchange); the stream-ordered path (3 launches per step) measured 3.1 ms"}
    # validation step of the reference trainer (trainer.py:517-529): forward + label-smoothed CE + masked accuracy
    lg = m.decoder(enc, tg)
    tgt_next = torch.roll(tg, -1, dims=1)
    ms_x = timed(lambda: pkg.metrics.cross_entropy_metrics(lg, tgt_next, 0, 0.1, sync=False), reps=10)
    row_bytes = B * MAX_LEN * 512 * 4
    out["validation_loss_accuracy"] = {
        "workload": "SURVEY 8f-4: CrossEntropyLoss(ignore_index, mean, label_smoothing 0.1) + masked_accuracy over "
                    "%d x %d x 512 fp32 logits" % (B, MAX_LEN),
        "value": round(B * MAX_LEN / ms_x * 1e3, 1), "unit": "tokens/s", "ms": round(ms_x, 4),
        "roofline": {"bound": "hbm", "achieved": round(row_bytes / ms_x / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                     "frac": round(row_bytes / ms_x / 1e6 / pk["hbm"], 4), "algorithmic_bytes": row_bytes}}
    return out


def reference_runner(params=None):
    """(step_fn(x) -> None, kind, label): the LIVE reference's own `Seq2SeqModel.inference` (encoder +
    `_greedy_search`, model/seq2seq.py:124-232, with its per-row `.item()` host syncs) from the travelling copy under
    oracle/_ref (oracle/make_ref.py) when it is present -- kind "reference" -- else the oracle port of the same path
    (the same ATen CPU kernels) -- kind "port"."""
    import torch
    import oracle
    from oracle import ref_shim
    torch.set_num_threads(os.cpu_count())
    p = params if params is not None else oracle.make_params(CFG, 0)
    if ref_shim.available():
        try:
            m = ref_shim.build_reference_model(CFG, p)
            return (lambda x: m.inference(x, START, END, max_length=MAX_LEN)), "reference", \
                "the reference's own Seq2SeqModel.inference (greedy), fp32, torch CPU"
        except Exception as e:
            print("live reference unavailable (%r): timing the oracle port" % (e,), file=sys.stderr)

    def port_step(x):
        enc = oracle.encoder(p, x, CFG)
        oracle.greedy_search(p, enc, START, END, MAX_LEN, 1.0, CFG)
    return port_step, "port", "oracle port of encoder + _greedy_search, fp32, torch CPU"


def cpu_baseline(model=None, sample_batch=32, budget_s=20.0, min_reps=1):
    """The reference's CPU implementation of the path timed on the host cores on a bounded sample: BASELINE configs[0]
    (batch 32, 150 greedy steps), repeated until ~budget_s of CPU work."""
    import torch
    p = {k: v.detach().float().cpu() for k, v in model.state_dict().items()} if model is not None else None
    step, kind, label = reference_runner(p)
    x = torch.randn(sample_batch, 3, 64, 320, generator=torch.Generator().manual_seed(7))
    times = []
    with torch.no_grad():
        step(x[:4])                                                                              # warm-up
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            step(x)
            times.append(time.perf_counter() - t0)
            if len(times) >= min_reps and time.perf_counter() - t_all > budget_s:
                break
    med = statistics.median(times)
    return {"value": round(sample_batch / med, 2), "unit": "images/s", "cores": torch.get_num_threads(),
            "kind": kind, "sample": "batch %d x %d greedy steps (BASELINE configs[0]), median of %d passes; %s"
                                    % (sample_batch, MAX_LEN, len(times), label)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the box's host cores, all threads, on OUR
    arm's workload -- every step is one full batch of `--batch` (1024) images through encoder + 150 greedy steps.
    Rank 0 only; the other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    step, kind, label = reference_runner()
    B = args.batch
    x = torch.randn(B, 3, 64, 320, generator=torch.Generator().manual_seed(7))
    W = max(args.warmup, 1)
    with torch.no_grad():
        # the driver passes our arm's --steps / --warmup; a CPU step takes seconds, so the warm-up runs on a slice of
        # the batch (it only has to fault in the weights and spin up the thread pool) and the timed steps are full
        for _ in range(W):
            step(x[:32])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(x)
        dt = time.perf_counter() - t0
    v = round(B * args.steps / dt, 2)
    sample = "each step = the full batch of %d images, encoder + %d greedy steps; %s" % (B, MAX_LEN, label)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": W, "ms_per_step": round(dt / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: CNN-LSTM greedy decode, batch %d per GPU, 3x64x320 images, "
                               "max_len 150, V=512, E=H=256, L=1, random init" % B,
                   "global_batch": B, "note": "CPU arm: rank 0 only, fp32 normalised images (the reference's input dtype)"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--input-dtype", default="uint8", choices=["uint8", "bf16", "fp32"],
                    help="element type of the image tensors (HBM-resident for `value`, pinned host for `e2e`)")
    ap.add_argument("--beam-batch", type=int, default=512, help="images per GPU for the beam-5 line")
    ap.add_argument("--no-extras", action="store_true", help="skip the beam-5 and other-host-dtype measurements")
    ap.add_argument("--xchg-mode", default="side", choices=["side", "inline", "none"],
                    help="diagnostic (N > 1): side = the p2p exchange on its own stream beside the next decode (default); "
                         "inline = read + write on the compute stream right after the decode; none = no exchange at all "
                         "(what 8 independent replicas cost on this box)")
    ap.add_argument("--nccl-gather", action="store_true",
                    help="N > 1: token all-gather through NCCL on the compute stream (the checked reference of the p2p exchange)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
