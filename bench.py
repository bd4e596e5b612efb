#!/usr/bin/env python
"""bench.py -- DEXTRACTOR compression hot path on B200: dexqv/undexqv (and dexta/undexta) GB/s of
uncompressed data, with the HBM roofline of the dominant kernel and the reference's CPU tools
timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size-gb G]

A "step" is one pass of the hot path over one batch: dexqv (statistics scan, code construction,
encode) of a synthetic .quiva shard followed by undexqv (decode) of the result.  Workload =
BASELINE.json configs[1]: a 2 GB RS II-like .quiva per GPU (weak scaling: every rank holds its
own 2 GB shard of one logical file; the ranks exchange only the histograms and a few integers).

value   uncompressed GB/s summed over both directions (2*U / t), inputs resident in HBM, timed
        with CUDA events on the library's stream, max over ranks.
e2e     the same step through the host-buffer C ABI (dx_dexqv_host / dx_undexqv_host): pinned
        host -> device copies of the inputs and device -> host copies of the results included.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GB = 1e9


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
#  reference arm: the reference's own CPU tools (oracle/_ref), single threaded like the reference
# ------------------------------------------------------------------------------------------------

_SAMPLES = {}


def cpu_reference_sample(sample_mb: float, seed: int = 1, nproc: int = 1):
    """dexqv + undexqv of a bounded sample with the reference binaries; returns timings.
    nproc > 1: that many independent copies at once (the reference has no threads; independent
    files are the only way it uses more cores), throughput = nproc * bytes / wall."""
    import numpy as np
    from dextractor_b200 import synth
    from oracle import orc
    key = (sample_mb, seed)
    if key not in _SAMPLES:                     # generating the text is not part of the timing
        rng = np.random.default_rng(seed)
        L = synth.lengths_for_bytes(rng, int(sample_mb * 1e6), 5.0)
        _SAMPLES[key] = synth.make_quiva(seed, L)
    text = _SAMPLES[key]
    if orc.have_ref():
        if nproc > 1:
            enc, t_enc = orc.ref_tool_parallel("dexqv", text, nproc)
            back, t_dec = orc.ref_tool_parallel("undexqv", enc, nproc)
        else:
            enc, t_enc = orc.ref_tool("dexqv", text, taskset=0)
            back, t_dec = orc.ref_tool("undexqv", enc, taskset=0)
        kind = "reference"
    else:                                   # the C restatement (oracle/dx_oracle.c)
        nproc = 1
        t0 = time.perf_counter(); enc = orc.dexqv(text); t_enc = time.perf_counter() - t0
        t0 = time.perf_counter(); back = orc.undexqv(enc); t_dec = time.perf_counter() - t0
        kind = "port"
    assert back == text
    return {"bytes": len(text) * nproc, "t_enc": t_enc, "t_dec": t_dec, "kind": kind,
            "compressed": len(enc) * nproc, "cores": nproc, "sample_bytes": len(text)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_mb = args.ref_sample_mb
    nproc = max(1, (os.cpu_count() or 1) if args.ref_cores <= 0 else args.ref_cores)
    times = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(sample_mb, seed=1, nproc=nproc)
        if i >= args.warmup:
            times.append(r)
    t = sum(x["t_enc"] + x["t_dec"] for x in times) / len(times)
    U = times[0]["bytes"]
    nproc = times[0]["cores"]
    val = 2 * U / t / GB
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, int(args.size_gb * GB)),
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": nproc, "kind": times[0]["kind"],
                         "sample": f"{nproc} x {times[0]['sample_bytes']/1e6:.0f} MB synthetic .quiva per "
                                   f"step, dexqv then undexqv (reference binaries, one single-threaded "
                                   f"process per core on independent copies, files on /dev/shm)",
                         "dexqv_gbs": U / (sum(x['t_enc'] for x in times) / len(times)) / GB,
                         "undexqv_gbs": U / (sum(x['t_dec'] for x in times) / len(times)) / GB,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


METRIC = "dexqv+undexqv uncompressed GB/s"


def workload_config(args, U):
    return {"workload": "BASELINE.json configs[1]: dexqv/undexqv on a synthetic 2 GB RS II-like "
                        ".quiva per GPU (5 QV streams)",
            "uncompressed_bytes_per_gpu": int(U), "shards": args.gpus,
            "l2": "inputs (2 GB) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
#  our arm
# ------------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-gb", type=float, default=2.0, help="uncompressed .quiva GB per GPU")
    ap.add_argument("--ref-sample-mb", type=float, default=64.0, help="per process")
    ap.add_argument("--ref-cores", type=int, default=0, help="reference arm processes (0 = all cores)")
    ap.add_argument("--cpu-sample-mb", type=float, default=64.0, help="per process")
    ap.add_argument("--no-extras", action="store_true", help="skip the dexta/undexta side numbers")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import dextractor_b200 as dx
    from dextractor_b200 import lib as dxl
    from dextractor_b200 import shards, synth_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries ONE JSON line: keep NCCL's own banner out of it ("NCCL version ..." goes to
        # stdout at NCCL_DEBUG=VERSION and =WARN; this image sets one of them)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=dev)

    ctx = dx.Context(local)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    hbm_peak, peak_src = peaks()

    # ---- workload: one 2 GB shard per rank --------------------------------------------------
    target = int(args.size_gb * GB)
    text, nent, npos = synth_torch.make_quiva_device(100 + rank, target, dev,
                                                     well_base=rank * 2_000_000)
    torch.cuda.synchronize()
    U = text.numel()
    prefix = bytes(text[:200].cpu().numpy().tobytes())
    prefix = prefix[: prefix.index(b"/", 1)]
    enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8, device=dev)
    back = torch.empty(U + 4096, dtype=torch.uint8, device=dev)
    state = {}

    def exchange_stats(st):
        """the only inter-GPU exchange of the path, one NCCL all-gather per step: every rank's
        6x256 histograms, its position / entry counts and the last well of its shard (the offset
        hand-off); the sums are formed locally.  -> (summed statistics, last well of rank-1)"""
        if world == 1:
            return st, 0
        if "xbuf" not in state:                      # pinned staging + device buffers, allocated once
            k = 6 * 256 + 3
            state["xbuf"] = (torch.empty(k, dtype=torch.int64).pin_memory(),
                             torch.empty(k, dtype=torch.int64, device=dev),
                             torch.empty(world * k, dtype=torch.int64, device=dev),
                             torch.empty(world * k, dtype=torch.int64).pin_memory())
        mine_h, mine_d, allt, all_h = state["xbuf"]
        mine_h.numpy()[:] = shards.pack_stats(st, state["last_well"])
        mine_d.copy_(mine_h, non_blocking=True)
        dist.all_gather_into_tensor(allt, mine_d)
        all_h.copy_(allt, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        # the run characters were fixed by the first shard (rank 0 resolves them in its first ~100 k
        # positions) and handed to the other ranks before the loop
        return shards.merge_stats(all_h.numpy(), rank, state["rc"])

    def carry_for_rank():
        """rank r > 0 counts run lengths with rank 0's run characters from its first entry on"""
        if world == 1 or rank == 0:
            return None
        return state["carry"]

    def step_device():
        """dexqv: scan -> (allreduce) -> code construction -> file header + encode.
        The image is laid out as [header][entries] from enc[0] (16-byte aligned)."""
        st = ctx.qv_scan_dev(text.data_ptr(), U, carry_for_rank())
        tot, lwell_in = exchange_stats(st)
        cd = dxl.make_coding(tot, False)
        hdr = b"\xaa\x55" + dxl.write_coding(cd, prefix)
        hl = len(hdr)
        ctx.h2d(enc.data_ptr(), hdr)
        body, lastw_out, offs = ctx.qv_encode_dev(text.data_ptr(), U, cd, False, lwell_in,
                                                  enc.data_ptr() + hl, enc.numel() - hl,
                                                  want_offsets=nent)
        state.update(hdr=hdr, img_len=hl + body, offs=offs + hl, lwell_in=lwell_in)
        return hl + body

    def decode_known():
        m = ctx.undexqv_dev(enc.data_ptr(), state["img_len"], False, back.data_ptr(), back.numel(),
                            entry_off=state["offs"], well_in=state["lwell_in"])
        state["out_len"] = m
        return m

    def decode_discover():
        return ctx.undexqv_dev(enc.data_ptr(), state["img_len"], False, back.data_ptr(),
                               back.numel(), well_in=state["lwell_in"])

    def full_step():
        step_device()
        decode_known()

    # rank > 0 needs rank 0's run characters before its scan: resolve once, outside the loop
    if world > 1:
        st0 = ctx.qv_scan_dev(text.data_ptr(), U, None)
        rc = torch.tensor([st0.delchar, st0.subchar], dtype=torch.int64, device=dev)
        dist.broadcast(rc, 0)
        c = dx.Carry()
        c.delchar, c.subchar, c.totchar = int(rc[0]), int(rc[1]), 200000
        state["carry"] = c
        state["rc"] = (int(rc[0]), int(rc[1]))
        # last well of this shard: parse the last header once
        tail = bytes(text[-400000:].cpu().numpy().tobytes())
        k = tail.rindex(b"\n" + prefix + b"/")          # a QV line may start with '@' too
        state["last_well"] = int(tail[k + 1:].split(b"/")[1])
    else:
        state["last_well"] = 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps (device-resident) ---------------------------------------
    for _ in range(args.warmup):
        full_step()
    # property check at full size: decode(encode(x)) == x
    ok = bool(torch.equal(back[: state["out_len"]], text)) and state["out_len"] == U
    if not ok:
        raise SystemExit("round trip at full size differs from the input")

    barrier()
    ctx.launch_count(reset=True)
    ctx.profile(True); ctx.profile_report()
    sampler = ClockSampler(local); sampler.start()
    with torch.cuda.stream(ext):
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record(ext)
        marks = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            full_step()
            marks.append(time.perf_counter() - t0)
        ev1.record(ext)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = ctx.launch_count()
    prof = ctx.profile_report(); ctx.profile(False)
    if world > 1:
        tms = torch.tensor([ms], device=dev); dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms[0])
        tb = torch.tensor([U, state["img_len"]], dtype=torch.int64, device=dev); dist.all_reduce(tb)
        U_all, C_all = int(tb[0]), int(tb[1])
    else:
        U_all, C_all = U, state["img_len"]
    ms_step = ms / args.steps
    value = 2 * U_all / (ms_step * 1e-3) / GB

    # per-direction timing (device resident), a few reps each
    def timed(fn, reps=3):
        best = []
        for _ in range(reps):
            barrier()
            with torch.cuda.stream(ext):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(ext); fn(); b.record(ext)
            barrier()
            best.append(a.elapsed_time(b))
        return min(best), sorted(best)[len(best) // 2]

    enc_ms, _ = timed(step_device)
    dec_known_ms, _ = timed(decode_known)
    decode_discover(); decode_discover()        # the scratch arena settles at this path's size
    dec_disc_ms, _ = timed(decode_discover)
    decode_known()
    C = state["img_len"]

    # ---- roofline of the dominant kernel (CUDA events inside the library, timed region) -------
    # algorithmic bytes per launch: hist 0.8U ; size U ; emit U+C ; decode C+U ; walk C
    lines_bytes = 5 * (npos + nent)                    # the 5 QV lines incl. newlines
    algo = {"k_qv_hist_plain": 0.4 * lines_bytes, "k_qv_hist_run": 0.4 * lines_bytes,
            "k_qv_size": lines_bytes, "k_qv_emit": lines_bytes + C,
            "k_qv_decode5": C + U, "k_qv_decode5_spec": C + lines_bytes, "k_qv_assemble": 2 * U,
            "k_pred_slots": U}
    top = max(prof.items(), key=lambda kv: kv[1][1]) if prof else ("none", (1, 1.0))
    tname, (tcalls, ttot) = top
    tavg = ttot / max(tcalls, 1)
    achieved = algo.get(tname, U) / (tavg * 1e-3) / GB
    traffic = None                              # dram bytes per launch from the committed ncu capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tname in tr["kernels"]:
            traffic = tr["kernels"][tname]["dram_bytes_per_text_byte"] * U
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": tname, "achieved": achieved, "peak": hbm_peak,
                "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                "peak_source": peak_src, "avg_ms": tavg,
                "share_of_step": ttot / max(sum(v[1] for v in prof.values()), 1e-9)}
    path = {"dexqv": {"algorithmic_bytes": 1.8 * U + C, "ms": enc_ms,
                      "frac": (1.8 * U + C) / (enc_ms * 1e-3) / GB / hbm_peak},
            "undexqv_offsets_known": {"algorithmic_bytes": C + U, "ms": dec_known_ms,
                                      "frac": (C + U) / (dec_known_ms * 1e-3) / GB / hbm_peak},
            "undexqv_offsets_discovered": {"algorithmic_bytes": 2 * C + U, "ms": dec_disc_ms,
                                           "frac": (2 * C + U) / (dec_disc_ms * 1e-3) / GB / hbm_peak}}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------
    h_text = torch.empty(U, dtype=torch.uint8).pin_memory()
    h_text.copy_(text)
    h_enc = torch.empty(U // 2 + (1 << 20), dtype=torch.uint8).pin_memory()
    h_back = torch.empty(U + 4096, dtype=torch.uint8).pin_memory()
    import ctypes as C_
    L = ctx.L

    def e2e_step():
        n1 = C_.c_size_t(0)
        ctx._check(L.dx_dexqv_host(ctx.h, h_text.data_ptr(), U, 0, h_enc.data_ptr(), h_enc.numel(),
                                   C_.byref(n1)))
        n2 = C_.c_size_t(0)
        ctx._check(L.dx_undexqv_host(ctx.h, h_enc.data_ptr(), n1.value, 0, h_back.data_ptr(),
                                     h_back.numel(), C_.byref(n2)))
        return n1.value, n2.value

    e2e = None
    if world == 1:
        n1, n2 = e2e_step()
        assert n2 == U and bool(torch.equal(h_back[:U], h_text))
        barrier()
        with torch.cuda.stream(ext):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(ext)
            for _ in range(args.steps):
                n1, n2 = e2e_step()
            b.record(ext)
        barrier()
        e2e_ms = a.elapsed_time(b) / args.steps
        e2e = {"value": 2 * U / (e2e_ms * 1e-3) / GB, "unit": "GB/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(U + n1), "d2h_bytes_per_step": int(n1 + n2),
               "api": "dx_dexqv_host + dx_undexqv_host (pinned host buffers; the decoder "
                      "rediscovers entry offsets from the file)"}
    else:
        # every rank runs its own shard end to end through the same host-buffer calls
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(ext):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(ext)
            for _ in range(args.steps):
                n1, n2 = e2e_step()
            b.record(ext)
        barrier()
        tms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms = float(tms[0])
        e2e = {"value": 2 * U_all / (e2e_ms * 1e-3) / GB, "unit": "GB/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int((U + n1) * world), "d2h_bytes_per_step": int((n1 + n2) * world),
               "api": "dx_dexqv_host + dx_undexqv_host per rank on its own shard (independent "
                      "per-shard files)"}
    del h_text, h_enc, h_back

    extras = {"dexqv_gbs": U / (enc_ms * 1e-3) / GB,
              "undexqv_offsets_known_gbs": U / (dec_known_ms * 1e-3) / GB,
              "undexqv_offsets_discovered_gbs": U / (dec_disc_ms * 1e-3) / GB,
              "compressed_bytes": int(C), "ratio": U / C, "entries": nent,
              "kernels_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(prof.items())}}

    # ---- side numbers: dexta/undexta on a 1 GB fasta (configs[0]) ---------------------------------
    if not args.no_extras and rank == 0:
        del back
        torch.cuda.empty_cache()
        fa, nfa = synth_torch.make_fasta_device(7, int(1.0 * GB), dev)
        UF = fa.numel()
        pk = torch.empty(UF // 3 + (1 << 20), dtype=torch.uint8, device=dev)
        un = torch.empty(UF + 4096, dtype=torch.uint8, device=dev)
        m = ctx.dexta_dev(dx.FASTA, fa.data_ptr(), UF, pk.data_ptr(), pk.numel())
        k = ctx.undexta_dev(dx.FASTA, pk.data_ptr(), m, 80, False, un.data_ptr(), un.numel())
        assert k == UF and bool(torch.equal(un[:UF], fa)), "dexta/undexta round trip differs"
        pack_ms, _ = timed(lambda: ctx.dexta_dev(dx.FASTA, fa.data_ptr(), UF, pk.data_ptr(), pk.numel()))
        unpack_ms, _ = timed(lambda: ctx.undexta_dev(dx.FASTA, pk.data_ptr(), m, 80, False,
                                                     un.data_ptr(), un.numel()))
        extras.update(dexta_gbs=UF / (pack_ms * 1e-3) / GB, undexta_gbs=UF / (unpack_ms * 1e-3) / GB,
                      dexta_frac=(UF + m) / (pack_ms * 1e-3) / GB / hbm_peak,
                      undexta_frac=(UF + m) / (unpack_ms * 1e-3) / GB / hbm_peak,
                      fasta_bytes=int(UF), dexta_bytes=int(m))
        del fa, pk, un
        torch.cuda.empty_cache()

        # configs[2]: dexar/undexar on a synthetic 1 GB Sequel-style .arrow
        ar, nar = synth_torch.make_fasta_device(9, int(1.0 * GB), dev, arrow=True)
        UA = ar.numel()
        pk = torch.empty(UA // 3 + (1 << 20), dtype=torch.uint8, device=dev)
        un = torch.empty(UA + 4096, dtype=torch.uint8, device=dev)
        m = ctx.dexta_dev(dx.ARROW, ar.data_ptr(), UA, pk.data_ptr(), pk.numel())
        k = ctx.undexta_dev(dx.ARROW, pk.data_ptr(), m, 80, False, un.data_ptr(), un.numel())
        # SN=%.2f headers pass through a float and lose up to 0.01 per trip, so the
        # property at full size is: same length, and only SNR digits of the header lines may differ
        ndiff = int((un[:UA] != ar).sum()) if k == UA else -1
        assert k == UA and 0 <= ndiff <= 16 * nar, "dexar/undexar round trip differs outside the SNR digits"
        par_ms, _ = timed(lambda: ctx.dexta_dev(dx.ARROW, ar.data_ptr(), UA, pk.data_ptr(), pk.numel()))
        unar_ms, _ = timed(lambda: ctx.undexta_dev(dx.ARROW, pk.data_ptr(), m, 80, False,
                                                   un.data_ptr(), un.numel()))
        extras.update(dexar_gbs=UA / (par_ms * 1e-3) / GB, undexar_gbs=UA / (unar_ms * 1e-3) / GB,
                      arrow_bytes=int(UA), dexar_bytes=int(m))
        del ar, pk, un
        torch.cuda.empty_cache()

        # configs[4]: mixed short/long subread lengths (500 bp - 50 kb), 0.5 GB each
        sweep = {}
        rs = np.random.default_rng(5)
        npos_t = int(0.5 * GB / 5.02)
        dists = {"log_uniform_500_50k": lambda k: np.exp(rs.uniform(np.log(500), np.log(50000), size=k)),
                 "90pct_500bp_10pct_50kb": lambda k: np.where(rs.random(k) < 0.9, 500, 50000),
                 "10pct_500bp_90pct_50kb": lambda k: np.where(rs.random(k) < 0.1, 500, 50000)}
        for name, draw in dists.items():
            Ls = np.asarray(draw(200000), dtype=np.int64)
            Ls = Ls[: int(np.searchsorted(np.cumsum(Ls), npos_t)) + 1]
            tx, ne, npz = synth_torch.make_quiva_device(50, 0, dev, lengths=Ls)
            Us = tx.numel()
            e2 = torch.empty(Us // 2 + (1 << 20), dtype=torch.uint8, device=dev)
            b2 = torch.empty(Us + 4096, dtype=torch.uint8, device=dev)
            st2 = {}

            def enc2():
                st2["n"] = ctx.dexqv_dev(tx.data_ptr(), Us, False, e2.data_ptr(), e2.numel())

            def dec2():
                st2["m"] = ctx.undexqv_dev(e2.data_ptr(), st2["n"], False, b2.data_ptr(), b2.numel())

            enc2(); dec2(); dec2()
            assert st2["m"] == Us and bool(torch.equal(b2[:Us], tx)), "length sweep round trip differs"
            t_e, _ = timed(enc2)
            t_d, _ = timed(dec2)
            sweep[name] = {"entries": int(ne), "bytes": int(Us), "dexqv_gbs": Us / (t_e * 1e-3) / GB,
                           "undexqv_discovered_gbs": Us / (t_d * 1e-3) / GB}
            del tx, e2, b2
            torch.cuda.empty_cache()
        extras["length_sweep"] = sweep

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r1 = cpu_reference_sample(args.cpu_sample_mb, seed=1, nproc=1)
        r = cpu_reference_sample(args.cpu_sample_mb, seed=1, nproc=os.cpu_count() or 1)
        t = r["t_enc"] + r["t_dec"]
        cpu = {"value": 2 * r["bytes"] / t / GB, "unit": "GB/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{r['cores']} x {r['sample_bytes']/1e6:.0f} MB synthetic .quiva (same generator "
                         f"family), dexqv then undexqv, reference binaries, one single-threaded process per "
                         f"core on independent copies, /dev/shm",
               "dexqv_gbs": r["bytes"] / r["t_enc"] / GB, "undexqv_gbs": r["bytes"] / r["t_dec"] / GB,
               "one_core": {"value": 2 * r1["bytes"] / (r1["t_enc"] + r1["t_dec"]) / GB,
                            "dexqv_gbs": r1["bytes"] / r1["t_enc"] / GB,
                            "undexqv_gbs": r1["bytes"] / r1["t_dec"] / GB},
               "host_cores_available": os.cpu_count()}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": workload_config(args, U),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "path_roofline": path, "cpu_baseline": cpu,
                "extra": extras}
        print(json.dumps(line), flush=True)
    # Teardown order matters: the pinned staging buffers of exchange_stats were used on the
    # library's stream (`ext`), and torch's pinned-memory allocator records an event on every
    # stream that used a block when the block is freed.  Freed after ctx.close() had destroyed
    # that stream, the record threw "CUDA error: context is destroyed" from a tensor destructor
    # and every rank of an N>1 run ended in SIGABRT after printing its line.  So: drop them first.
    state.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
