#!/usr/bin/env python
"""bench.py — the hot-path benchmark (BASELINE.json metric: IQ samples/sec/train-step, DGRU, APA_200MHz frames).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload c2a] [--settle 5000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the reference's net_train body (modules/train_funcs.py:33-48): forward + I/Q MSE + backward
+ clip_grad_norm_(200) + AdamW + read-back of the loss, on one batch of REAL measured frames: the shipped APA_200MHz training
stream (tests/golden/iq_streams.npz, made by oracle/make_iq_streams.py through the reference's own loader), framed exactly like
IQFrameDataset (stride-1 windows, modules/data_collector.py:233-252) in the order of the seeded epoch permutation.
Workload at N=1 = BASELINE.json configs[1] (C2a in BASELINE.md): DGRU H=13 (1041 params), batch 64, frame length 2048, fp32,
train_pa.  N>1: every rank takes its 64-frame shard of a 64*N global batch (weak scaling), one exchange of the flat
1041-float gradient (+loss) per step.  Before anything is timed the model is TRAINED for --settle (default 5000) real steps,
so the time-chunk plan that is measured is the one trained weights need; the all-serial kernels are timed too.
Prints ONE JSON line (see DESIGN.md §5 for every key)."""
import argparse, copy, json, os, sys, threading, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.md §2 configs; c2a is the one the metric is quoted on (configs[1]) and the default.  B = global batch of the config.
    "c1": dict(name="C1: GRU H=32 (3522 params) train_pa step, B=8 x T=1024, fp32", kind="gru", H=32, B=8, T=1024, dataset="DPA_100MHz"),
    "c2a": dict(name="C2a: DGRU H=13 (1041 params) train_pa step, B=64 x T=2048 IQ frames, fp32", kind="dgru", H=13, B=64, T=2048,
                dataset="APA_200MHz"),
    "c2b": dict(name="C2b: DGRU H=13 DPD -> frozen DGRU H=13 PA, train_dpd step, B=64 x T=2048, fp32", kind="dgru", H=13, B=64, T=2048,
                pa=("dgru", 13), dataset="APA_200MHz"),
    "c3": dict(name="C3: TRes-DeltaGRU H=15 (999 params, thx .01 thh .05) DPD -> frozen DGRU H=23 PA, B=256 x T=2048, fp32",
               kind="deltagru_tcnskip", H=15, B=256, T=2048, pa=("dgru", 23), dataset="APA_200MHz"),
    "c3b": dict(name="C3 with bf16 IQ storage (BASELINE configs[2]): TRes-DeltaGRU H=15 DPD -> frozen DGRU H=23 PA, B=256 x T=2048, bf16 in HBM / fp32 arithmetic",
                kind="deltagru_tcnskip", H=15, B=256, T=2048, pa=("dgru", 23), io="bf16", dataset="APA_200MHz"),
    "c3s": dict(name="C3 (script frame length): TRes-DeltaGRU H=15 DPD -> frozen DGRU H=23 PA, B=256 x T=200, fp32",
                kind="deltagru_tcnskip", H=15, B=256, T=200, pa=("dgru", 23), dataset="APA_200MHz"),
    "c4p": dict(name="C4: PGJANET H=15 (1727 params) step, global B=1024 x T=4096, fp32", kind="pgjanet", H=15, B=1024, T=4096,
                dataset="APA_200MHz_b"),
    "c4d": dict(name="C4: DVRJANET H=15 K=3 (1685 params) step, global B=1024 x T=4096, fp32", kind="dvrjanet", H=15, B=1024, T=4096,
                dataset="APA_200MHz_b"),
    "c5g": dict(name="C5: GMP (495 params) step, global B=512 x T=50, fp32", kind="gmp", H=1, B=512, T=50, dataset="DPA_200MHz"),
    "c5q": dict(name="C5: QGRU H=10 W8A8 QAT (515 params) step, global B=512 x T=50, fp32 fake-quant", kind="qgru_qat", H=10, B=512, T=50,
                dataset="DPA_200MHz"),
    "lstm": dict(name="LSTM H=9 (488 params) train_pa step, B=64 x T=2048, fp32", kind="lstm", H=9, B=64, T=2048, dataset="APA_200MHz"),
    # SURVEY §8 row f-4 (every other backbone models.py can build) and the layered path, train_pa step, B=64 x T=2048 real frames
    "f4_vdlstm": dict(name="VDLSTM H=9", kind="vdlstm", H=9, B=64, T=2048, dataset="APA_200MHz"),
    "f4_bojanet": dict(name="BOJANET H=10", kind="bojanet", H=10, B=64, T=2048, dataset="APA_200MHz"),
    "f4_apnrru": dict(name="APNRRU H=8", kind="apnrru", H=8, B=64, T=2048, dataset="APA_200MHz"),
    "f4_deltajanet": dict(name="DeltaJANET H=10", kind="deltajanet", H=10, B=64, T=2048, dataset="APA_200MHz"),
    "f4_mcldnn": dict(name="MCLDNN C=8", kind="mcldnn", H=8, B=64, T=2048, dataset="APA_200MHz"),
    "f4_rvtdcnn": dict(name="RVTDCNN H=6", kind="rvtdcnn", H=6, B=64, T=2048, dataset="APA_200MHz"),
    "f4_tcnn": dict(name="TCNN C=8", kind="tcnn", H=8, B=64, T=2048, dataset="APA_200MHz"),
    "f4_neuraltx": dict(name="NeuralTX C=8", kind="neuraltx", H=8, B=64, T=2048, dataset="APA_200MHz"),
    "qat_tres": dict(name="fake-quantised W16A16 TRes-DeltaGRU H=15 DPD -> frozen DGRU H=23 PA, B=64 x T=200 (OpenDPDv2.sh QAT stage)", kind="tres_qat",
                     H=15, B=64, T=200, pa=("dgru", 23), dataset="APA_200MHz"),
    "wide_dgru64": dict(name="DGRU H=64 (layered path)", kind="dgru", H=64, B=64, T=2048, dataset="APA_200MHz"),
    "wide_gru32x2": dict(name="GRU H=32, 2 layers (layered path)", kind="gru", H=32, L=2, B=64, T=2048, dataset="APA_200MHz"),
}
OTHER_BACKBONES = ("qat_tres", "f4_vdlstm", "f4_bojanet", "f4_apnrru", "f4_deltajanet", "f4_mcldnn", "f4_rvtdcnn", "f4_tcnn", "f4_neuraltx", "wide_dgru64",
                   "wide_gru32x2")   # 1-GPU lines only: coverage, ms per whole train step
SECONDARY = ("c2a", "c3", "c3b", "c4p", "c4d", "c5g", "c5q")   # extra keys of the line: fixed GLOBAL batch split over the ranks (strong scaling)
ALGO_BYTES_PER_SAMPLE_PER_KERNEL = 16  # fwd reads x(8)+target(8); bwd re-reads x(8)+target/dout(8)  => 32 B/sample/step
STREAMS = os.path.join(ROOT, "tests", "golden", "iq_streams.npz")


def config_of(wl, world, weak):
    """The `config` object of the JSON line — a pure function of (workload, N) so that both arms print the same one."""
    gb = wl["B"] * world if weak else wl["B"]
    return {"workload": wl["name"], "dataset": wl["dataset"] + " train stream (real measured frames, stride-1 windows)",
            "global_batch": gb, "per_gpu_batch": gb // world, "frame_len": wl["T"], "parallelism": f"dp{world}",
            "iq_storage": wl.get("io", "f32"), "optimizer": "clip_grad_norm_(200)+AdamW(lr=5e-4)",
            "l2": "inputs larger than L2: the timed steps walk a pool of distinct materialised batches of real frames (>= 160 MB where the "
                  "batch size allows, see run.pool_mb), read in place by the kernels; L2 flushed (256 MiB write) after warm-up"}


# ------------------------------------------------------------------------------------------------ data
class Feed:
    """Real frames of one dataset for one rank.  The raw (N,2) streams live on the device (and the host); batch `s` of the run is
    frames perm_e[k*GB + rank*B : ... + B] of epoch e = s // batches_per_epoch (the seeded permutation the reference's
    DataLoader(shuffle=True) draws, dp.epoch_permutation / dp.shard_batch_indices), as frame START indices."""

    def __init__(self, wl, dev, rank, world, B, GB):
        import torch
        z = np.load(STREAMS)
        ds, self.T, self.B, self.GB, self.rank = wl["dataset"], wl["T"], B, GB, rank
        x = torch.from_numpy(z[ds + ".x"])
        y = torch.from_numpy(z[ds + (".dpd_target" if "pa" in wl else ".y")])     # train_dpd: y = float32(gain * x_f64) (project.py:221-225)
        if wl.get("io") == "bf16":                                                # storage only: the kernels widen every sample exactly
            x, y = x.bfloat16(), y.bfloat16()
        self.x_host, self.y_host = x.contiguous(), y.contiguous()
        self.x = self.y = None
        if dev is not None:
            self.x, self.y = self.x_host.to(dev), self.y_host.to(dev)
        self.n_frames = x.shape[0] - self.T + 1
        self.per_epoch = max(1, self.n_frames // GB)
        self._perm = {}

    def table(self, first, n):
        """int32 (n, B) host tensor: frame starts of this rank's shard of global batches first .. first+n-1."""
        import torch
        from opendpd_b200 import dp
        out = torch.empty(n, self.B, dtype=torch.int32)
        for i in range(n):
            e, k = divmod(first + i, self.per_epoch)
            if e not in self._perm:
                self._perm = {e: dp.epoch_permutation(self.n_frames, seed=0, epoch=e)}
            idx, _ = dp.shard_batch_indices(self._perm[e], k, self.GB, self.rank, self.GB // self.B)
            out[i] = idx.to(torch.int32)
        return out

    def frames(self, starts):
        """Materialised (n,B,T,2) host frames for a (n,B) table of starts — what IQFrameDataset would hand the DataLoader."""
        import torch
        idx = starts.long()[..., None] + torch.arange(self.T)[None, None, :]
        return self.x_host[idx].contiguous(), self.y_host[idx].contiguous()


class ClockSampler:
    """SM clock / throttle-reason sampler (NVML, the library behind nvidia-smi) running during the timed regions."""

    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag, self.thread, self.ok = [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def _loop(self):
        nv = self.nv
        names = {getattr(nv, k): k for k in dir(nv) if k.startswith("nvmlClocksThrottleReason") or k.startswith("nvmlClocksEventReason")}
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((clk, util))
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) and bit & (bit - 1) == 0:
                        self.reasons.add(name.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        clk = [c for c, _ in self.samples]
        bad = {"GpuIdle", "None", "All", "ApplicationsClocksSetting"}
        return {"sm_mhz": float(np.median(clk)), "sm_max_mhz": float(self.max_sm), "samples": len(clk),
                "reasons": sorted(r for r in self.reasons if r not in bad)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_leg(wl, B, seconds=10.0, threads=None):
    """CPU arm: the net_train body of the workload on the host cores, on the first real batch of the run.
    Two restatements are timed: the PyTorch-ATen op sequence the reference executes (oracle/torch_port.py — the reference IS
    PyTorch; this is what its CPU path costs) and the plain-C/OpenMP port (oracle/odpd_oracle.c, forward+MSE+backward)."""
    import torch
    from oracle import oracle
    kind, H, T = wl["kind"], wl["H"], wl["T"]
    if kind in ("qgru_qat", "tres_qat"):
        return {}          # QAT flows: no single-call CPU arm
    cores = threads or os.cpu_count() or 1
    feed = Feed(wl, None, 0, 1, B, B)
    xs, ys = feed.frames(feed.table(0, 1))
    xt, yt = xs[0].float(), ys[0].float()
    res = {}
    if "pa" not in wl:     # the C port is one backbone per call; the cascade is timed on the PyTorch-op restatement only
        x, y = xt.numpy(), yt.numpy()
        rng = np.random.default_rng(0)
        P = oracle.n_params(kind, H)
        params = (0.3 * rng.standard_normal(P)).astype(np.float32)
        nthr = min(cores, B)
        oracle.run(kind, x, params, target=y, H=H, nthreads=nthr)
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds / 2 or n < 3:
            oracle.run(kind, x, params, target=y, H=H, nthreads=nthr)
            n += 1
        c_dt = (time.perf_counter() - t0) / n
        res["c_port"] = {"value": B * T / c_dt, "unit": "IQ samples/s", "cores": nthr, "s_per_step": c_dt,
                         "sample": f"{n} x (fwd+MSE+bwd) of the full {B}x{T} batch of real frames, C/OpenMP over sequences"}
    try:
        from oracle import torch_port
        step = torch_port.make_train_step(kind, H, seed=0, thx=0.01, thh=0.05, pa=wl.get("pa"))
        # PyTorch's intra-op pool degrades badly when oversubscribed on these ~1e3-element ops: use all host threads it can
        # USE — scan a few pool sizes (one step each) and keep the fastest; the count is reported in `cores`.
        best = None
        for nthr_t in sorted({cores, min(cores, 32), min(cores, 16), min(cores, 8), min(cores, 4)}):
            torch.set_num_threads(nthr_t)
            t1 = time.perf_counter(); step(xt, yt); d1 = time.perf_counter() - t1
            if best is None or d1 < best[1]:
                best = (nthr_t, d1)
            elif d1 > 2.0 * best[1]:
                break                     # larger pools only get slower from here (measured: 128 threads = 17x slower than 4)
        torch.set_num_threads(best[0])
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds / 2 or n < (3 if "pa" not in wl else 2):
            step(xt, yt)
            n += 1
        dt = (time.perf_counter() - t0) / n
        res["torch_port"] = {"value": B * T / dt, "unit": "IQ samples/s", "cores": best[0], "host_cores": cores, "s_per_step": dt,
                             "sample": f"{n} x full net_train body (fwd, MSE, bwd, clip 200, AdamW, loss.item()) of the {B}x{T} batch of real "
                                       f"{wl['dataset']} frames, PyTorch CPU ops (the reference's own op sequence), torch.set_num_threads({best[0]}) "
                                       f"= fastest of a scan up to {cores}"}
    except Exception as e:  # torch port optional
        res["torch_port_error"] = repr(e)
    return res


# ------------------------------------------------------------------------------------------------ stock-PyTorch GPU comparator
def gpu_reference_leg(wl, B, dev, n_warm=3, n_timed=10):
    """BASELINE.md §2 second comparator — the reference's op sequence executed by STOCK PyTorch on the same B200 (cuDNN GRU behind
    torch._VF.gru for gru/dgru/qgru/lstm, backbones/dgru.py:70; one ATen launch per Python op for the hand-written cells), full
    net_train body incl. loss.item(), on the first real batch.  Comparator only (like cpu_baseline): nothing of it ships."""
    import torch
    from oracle import torch_port, oracle
    kind, H, T = wl["kind"], wl["H"], wl["T"]
    feed = Feed(wl, None, 0, 1, B, B)
    xs, ys = feed.frames(feed.table(0, 1))
    x, y = xs[0].float().to(dev), ys[0].float().to(dev)
    g = torch.Generator().manual_seed(0)
    flat = torch.nn.Parameter((0.3 * torch.randn(oracle.n_params(kind, H), generator=g)).to(dev))
    pa_flat = None
    if "pa" in wl:
        pa_flat = (0.3 * torch.randn(oracle.n_params(wl["pa"][0], wl["pa"][1]), generator=g)).to(dev)
    opt = torch.optim.AdamW([flat], lr=5e-4)
    crit = torch.nn.MSELoss()

    def step():
        opt.zero_grad()
        out = torch_port.forward(kind, x, flat, H, 3, 0.01, 0.05)
        if pa_flat is not None:
            out = torch_port.forward(wl["pa"][0], out, pa_flat, wl["pa"][1])
        loss = crit(out, y)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([flat], 200.0)
        opt.step()
        return loss.item()
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n_timed):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    return {"ms_per_step": ms, "value": B * T / (ms * 1e-3), "unit": "IQ samples/s", "protocol": f"{n_warm} warm-up + {n_timed} timed steps, median",
            "what": "stock PyTorch " + torch.__version__ + " on the same GPU: " +
                    ("cuDNN RNN (torch._VF.gru) + ATen linear/relu/cat" if kind in ("gru", "dgru", "qgru", "qgru_amp1", "lstm") else "per-op ATen launches of the Python-loop cell") +
                    ", nn.MSELoss, clip_grad_norm_, torch AdamW, loss.item()"}


_REAL_STDOUT = None


def _quiet_stdout():
    """The driver parses stdout for ONE JSON line: route everything else (the reference-style 'Backbone Initialized...' print, NCCL's
    version banner, ...) to stderr at the file-descriptor level and keep the real stdout for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# ------------------------------------------------------------------------------------------------ native arm
def build_trainer(wl, dev, pg, world):
    import torch
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    torch.manual_seed(0)                               # same initial weights on every rank
    if wl["kind"] in ("qgru_qat", "tres_qat"):
        from opendpd_b200.quant import get_quant_model
        bits = 8 if wl["kind"] == "qgru_qat" else 16

        class _Proj:
            quant, n_bits_w, n_bits_a, pretrained_model = True, bits, bits, ""
        base = models.CoreModel(2, wl["H"], 1, "qgru") if wl["kind"] == "qgru_qat" else models.CoreModel(2, wl["H"], 1, "deltagru_tcnskip", thx=0.01, thh=0.05)
        net = get_quant_model(_Proj(), base).to(dev).train()
    else:
        net = models.CoreModel(2, wl["H"], wl.get("L", 1), wl["kind"], num_dvr_units=3, thx=0.01, thh=0.05).to(dev)
    if "pa" in wl:                                      # train_dpd: DPD in front of a frozen PA (steps/train_dpd.py:60-63)
        torch.manual_seed(1)
        pa_net = models.CoreModel(2, wl["pa"][1], 1, wl["pa"][0]).to(dev)
        net = models.CascadedModel(net, pa_net)
        net.freeze_pa_model()
    return net, NativeTrainStep(net, lr=5e-4, grad_clip_val=200.0, process_group=pg, world_size=world)


def chunk_report(trainer, B, T):
    from opendpd_b200.functional import chunk_reruns, chunk_worst_mismatch
    info = []
    launches = 1 if (trainer.px is None and trainer.world == 1 and trainer.fuse_optimizer) else 2     # reduce_partials (+ clip_adamw in its last CTA) | + optimiser launch
    for mod, backward, bi, save, need_dw in trainer.chunk_calls():
        sp_ = mod._spec()
        plan = sp_.chunk_plan(B, T, backward, save, need_dw)
        buf = trainer._bufs[bi].get("ws" if backward else "saved")
        info.append({"cell": mod.cell, "dir": "bwd" if backward else "fwd", "chunks": plan[0], "steps_per_chunk": plan[1],
                     "warmup_steps": plan[2], "serial_reruns": chunk_reruns(sp_, buf, B, T, backward, save, need_dw),
                     "worst_boundary_mismatch_over_tolerance": chunk_worst_mismatch(sp_, buf, B, T, backward, save, need_dw)})
        launches += 2 if plan[0] > 1 else 1
    return info, launches


class Run:
    """One workload on this rank: trainer + real-data feed + a pool of materialised batches larger than L2."""

    def __init__(self, wl, dev, pg, world, rank, weak, settle):
        import torch
        self.torch, self.wl, self.dev, self.pg, self.world, self.rank = torch, wl, dev, pg, world, rank
        GB = wl["B"] * world if weak else wl["B"]
        assert GB % world == 0, (GB, world)
        self.GB, self.B, self.T = GB, GB // world, wl["T"]
        self.multi = int(os.environ.get("ODPD_BENCH_STEPS_PER_REPLAY", "8"))
        B, T = self.B, self.T
        self.net, self.trainer = build_trainer(wl, dev, pg, world)
        self.feed = Feed(wl, dev, rank, world, B, GB)
        self.sink = torch.zeros(1, dtype=torch.float64, device=dev)
        # ---- untimed: TRAIN on the real stream (on-device framing, the seeded epoch order) so that what is timed afterwards is
        # the plan trained weights need, not the one a fresh initialisation allows
        self.settle = 0
        done = 0
        while done < settle:
            n = min(1000, settle - done)
            tab = self.feed.table(done, n).to(dev)
            for i in range(n):
                self.trainer.step_indexed(self.feed.x, self.feed.y, tab[i], T, loss_out=self.sink)
            done += n
        self.settle = done
        # ---- pool of distinct materialised batches of real frames, larger than L2 where the batch size allows
        bytes_per_batch = 2 * B * T * 2 * self.feed.x_host.element_size()
        self.POOL = int(max(4, min(80, (160 << 20) // bytes_per_batch + 1)))
        ptab = self.feed.table(self.settle, self.POOL)
        xs, ys = self.feed.frames(ptab)
        self.xs_pin, self.ys_pin = xs.pin_memory(), ys.pin_memory()
        self.xd, self.yd = xs.to(dev), ys.to(dev)
        self.px, self.py = self.xd.view(-1, 2), self.yd.view(-1, 2)
        self.ptable = ((torch.arange(self.POOL)[:, None] * B + torch.arange(B)[None, :]) * T).to(torch.int32).to(dev)
        self.pool_mb = self.POOL * bytes_per_batch / 2**20

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world > 1:
            t = self.torch.tensor([float(v)], device=self.dev)
            self.torch.distributed.all_reduce(t, op=self.torch.distributed.ReduceOp.MAX)
            return float(t.item())
        return float(v)

    def pool_step(self, i, loss_out):
        return self.trainer.step_indexed(self.px, self.py, self.ptable[i % self.POOL], self.T, loss_out=loss_out)

    def timed(self, W, K, flush=None, multi=None):
        """W warm-up + EXACTLY K timed steps over the pool, every step's loss copied device->host (pinned) inside the region;
        CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  Returns (ms total, losses).
        The steps are issued `multi` per CUDA-graph replay (NativeTrainStep.steps_indexed; 1 = one graph per step): one copy of the
        frame starts in and one copy of the losses out per replay instead of per step, no launch gap between the steps inside."""
        torch = self.torch
        multi = self.multi if multi is None else multi
        loss_pin = torch.zeros(K, dtype=torch.float64).pin_memory()
        order = (torch.arange(W + K) % self.POOL).to(self.dev)
        tab = self.ptable[order].contiguous()              # (W+K, B) frame starts of the warm-up and the timed steps, in order

        def issue(lo, n, out):
            if multi <= 1:
                for i in range(lo, lo + n):
                    self.trainer.step_indexed(self.px, self.py, tab[i], self.T, loss_out=None if out is None else out[i - W:i - W + 1])
            else:
                i = lo
                while i < lo + n:
                    m = min(multi, lo + n - i)
                    self.trainer.steps_indexed(self.px, self.py, tab[i:i + m], self.T, losses_out=None if out is None else out[i - W:i - W + m])
                    i += m
        # untimed: every replay shape the timed loop will use is run eagerly once, captured once and replayed once before the clock starts
        if multi > 1:
            for m in sorted({min(multi, K), K % multi} - {0}):
                for _ in range(3):
                    self.trainer.steps_indexed(self.px, self.py, tab[:m], self.T)
        issue(0, W, None)
        if flush is not None:
            flush.zero_()                               # evict the pool from L2: every timed step reads a cold batch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        issue(W, K, loss_pin)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), loss_pin.numpy().copy()

    def settle_plan(self, max_passes=12):
        """Let the chunk controller adapt to the pool batches: passes until a whole pass changed no plan (collectively)."""
        torch = self.torch
        n = 0
        for _ in range(max_passes):
            n_ev = len(self.trainer.chunk_events)
            for i in range(self.POOL + 8):
                self.pool_step(i, self.sink)
            n += self.POOL + 8
            if self.max_over_ranks(float(n_ev != len(self.trainer.chunk_events))) == 0.0:
                break
        return n

    def close(self):
        self.trainer._check_exchange(wait=True)
        if self.trainer.px is not None:
            self.torch.cuda.synchronize()
            if self.world > 1:
                self.torch.distributed.barrier()
            self.trainer.px.close()
            self.trainer.px = None


def dp_check(dev, pg, world, rank):
    """Data-parallel correctness, asserted on every multi-GPU run: 3 DP steps on rank shards == 3 single-process steps on the
    full global batch (same seeded permutation), and all replicas hold bit-identical parameters."""
    import torch
    import torch.distributed as dist
    from opendpd_b200 import models, dp
    from opendpd_b200.train import NativeTrainStep
    torch.manual_seed(0)
    net = models.CoreModel(2, 13, 1, "dgru").to(dev)
    ref = copy.deepcopy(net)
    z = np.load(STREAMS)
    stream = torch.from_numpy(z["APA_200MHz.x"][:6000]).to(dev)
    target = torch.from_numpy(z["APA_200MHz.y"][:6000]).to(dev)
    T, GB = 256, 8 * world
    perm = dp.epoch_permutation(stream.shape[0] - T + 1, seed=0)
    tr, trr = NativeTrainStep(net, process_group=pg, world_size=world), NativeTrainStep(ref)
    worst = 0.0
    for step in range(3):
        idx, n_global = dp.shard_batch_indices(perm, step, GB, rank, world)
        tr.step(dp.gather_frames(stream, idx, T), dp.gather_frames(target, idx, T), global_count=2 * n_global * T)
        ia = perm[step * GB:(step + 1) * GB]
        trr.step(dp.gather_frames(stream, ia, T), dp.gather_frames(target, ia, T))
        pa = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        pb = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
        worst = max(worst, float((pa - pb).abs().max().item()))
    allp = [torch.empty_like(pa) for _ in range(world)]
    dist.all_gather(allp, pa)
    same = all(torch.equal(allp[0], q) for q in allp)
    tr._check_exchange(wait=True)
    if tr.px is not None:
        torch.cuda.synchronize(); dist.barrier(); tr.px.close(); tr.px = None
    return {"replicas_identical": bool(same), "max_param_diff_vs_single": worst, "steps": 3,
            "exchange": "fused NVLink push (odpd_dp_clip_adamw)" if os.environ.get("ODPD_DP_P2P", "1") != "0" else "NCCL all-reduce"}


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--workload", default="c2a", choices=sorted(WORKLOADS), help="default c2a = BASELINE.json configs[1]")
    ap.add_argument("--settle", type=int, default=5000, help="real training steps before anything is timed (headline workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / gpu_reference legs")
    ap.add_argument("--no-secondary", action="store_true", help="skip the extra BASELINE.json configs (c3, c4, c5, strong-scaling c2a)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    wl = WORKLOADS[args.workload]
    T = wl["T"]
    n_decl = max(args.gpus, world)

    if args.impl == "reference":
        if rank != 0:
            return
        GB = wl["B"] * n_decl                     # the native arm's global batch (weak scaling): same config on both arms
        r = cpu_port_leg(wl, GB, seconds=max(args.cpu_seconds, 6.0))
        main_leg = r.get("torch_port") or r.get("c_port")
        if main_leg is None:
            emit({"impl": "reference", "unavailable": f"no CPU arm for workload {args.workload}: {r.get('torch_port_error', 'QAT flow')}"})
            return
        line = {"impl": "reference", "metric": "IQ samples/sec/train-step", "value": main_leg["value"], "unit": "IQ samples/s",
                "n_gpus": n_decl, "steps": K, "warmup": W, "ms_per_step": main_leg["s_per_step"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": wl["dataset"],
                "config": config_of(wl, n_decl, True),
                "cpu_baseline": {"value": main_leg["value"], "unit": "IQ samples/s", "cores": main_leg["cores"], "kind": "port",
                                 "sample": main_leg["sample"]},
                "cpu_c_port": r.get("c_port"),
                "e2e": {"value": main_leg["value"], "unit": "IQ samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if "torch_port_error" in r:
            line["torch_port_error"] = r["torch_port_error"]
        emit(line)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the native path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    dpc = dp_check(dev, pg, world, rank) if world > 1 else None

    run = Run(wl, dev, pg, world, rank, weak=True, settle=args.settle)
    trainer, B = run.trainer, run.B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    plan_settle = run.settle_plan()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- headline: K timed steps, device-resident real frames (pool > L2), loss read back every step
    ms, losses = run.timed(W, K, flush)
    assert np.all(np.isfinite(losses)), "non-finite loss in the timed region"
    chunk_info, launches = chunk_report(trainer, B, T)
    events_at_timed = list(trainer.chunk_events)
    # strict variant of the metric's definition: loss.item() (a host sync) after every step (train_funcs.py:48)
    run.barrier()
    ks = min(K, 100)
    t0 = time.perf_counter()
    for i in range(ks):
        float(run.pool_step(W + i, None).item())
    torch.cuda.synchronize()
    sync_ms = run.max_over_ranks((time.perf_counter() - t0) / ks * 1e3)

    # ---- e2e: host (pinned) frames, H2D inside the step, loss read back every step
    xs_pin, ys_pin, POOL = run.xs_pin, run.ys_pin, run.POOL
    for _ in range(3):      # untimed: the same block pattern as the timed call, so that every replay shape is run eagerly, captured and replayed once
        trainer.run_host_batches(((xs_pin[i % POOL], ys_pin[i % POOL]) for i in range(K)), steps_per_replay=max(run.multi, 1))
    run.barrier()
    t0 = time.perf_counter()
    e2e_losses = trainer.run_host_batches(((xs_pin[(W + i) % POOL], ys_pin[(W + i) % POOL]) for i in range(K)), steps_per_replay=max(run.multi, 1))
    torch.cuda.synchronize()
    e2e_s = run.max_over_ranks(time.perf_counter() - t0)
    assert len(e2e_losses) == K and all(np.isfinite(e2e_losses))
    run.barrier()
    t1 = time.perf_counter()
    for i in range(ks):
        trainer.step_host(xs_pin[(W + i) % POOL], ys_pin[(W + i) % POOL])
    torch.cuda.synchronize()
    e2e_seq_ms = run.max_over_ranks((time.perf_counter() - t1) / ks * 1e3)
    # on-device framing (SURVEY f-2), the way a real epoch runs: the raw APA stream resident in HBM (472 KB: L2-resident by nature),
    # per step the host sends only the B frame start indices of the seeded permutation, the loss comes back every step
    starts_pin = run.feed.table(run.settle + POOL, K + 8).pin_memory()
    ev_l = [torch.cuda.Event() for _ in range(2)]

    MULTI = run.multi
    loss_pin = [torch.zeros(max(MULTI, 1), dtype=torch.float64).pin_memory() for _ in range(2)]

    def indexed_steps(n, off):
        """n steps, MULTI per graph replay: per replay the (m,B) frame starts go pinned host -> device and the m losses come back; the
        host reads them one replay later."""
        pending, i, sl = None, 0, 0
        while i < n:
            m = min(max(MULTI, 1), n - i)
            if MULTI > 1:
                trainer.steps_indexed(run.feed.x, run.feed.y, starts_pin[off + i:off + i + m], T, losses_out=loss_pin[sl][:m])
            else:
                trainer.step_indexed(run.feed.x, run.feed.y, starts_pin[off + i], T, loss_out=loss_pin[sl][:1])
            ev_l[sl].record()
            if pending is not None:
                ev_l[pending[0]].synchronize()
                assert np.all(np.isfinite(loss_pin[pending[0]][:pending[1]].numpy()))
            pending = (sl, m)
            sl ^= 1
            i += m
        ev_l[pending[0]].synchronize()
    for _ in range(3):
        indexed_steps(min(K, max(MULTI, 1)), 0)          # eager, capture, replay of the replay shape
    if K % max(MULTI, 1):
        for _ in range(3):
            indexed_steps(K % MULTI, 0)
    run.barrier()
    t2 = time.perf_counter()
    indexed_steps(K, 3)
    torch.cuda.synchronize()
    e2e_idx_s = run.max_over_ranks(time.perf_counter() - t2)

    # ---- per-kernel durations for the roofline of the dominant kernel: CUDA events around a CUDA-graph replay of NK back-to-back
    # launches of the forward (distinct cold batches of the pool) and of the backward (on the batch the last forward saved, as in a
    # real step) — a replay keeps the device fed; issuing each launch from Python (~0.1 ms of host time) would time the host instead
    from opendpd_b200.functional import backbone_forward_raw, backbone_backward_raw
    bb = trainer.train_bb if "pa" not in wl else trainer.pa
    flat, _ = bb._flat_sync()
    spec = bb._spec()
    xd, yd = run.xd, run.yd
    NK = min(POOL, 64)
    count = float(2 * B * T)
    gflat = torch.empty_like(flat)
    kb0, kb1 = {}, {}
    out, _, saved = backbone_forward_raw(spec, xd[0], flat, yd[0], 1.0 / count, True, None, kb0)       # eager: allocates the buffers
    backbone_backward_raw(spec, xd[0], flat, saved, False, True, out=out, target=yd[0], gscale=2.0 / count, gflat=gflat, bufs=kb1)
    torch.cuda.synchronize()

    def replay_ms(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for i in range(NK):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / NK
    fwd_ms = replay_ms(lambda i: backbone_forward_raw(spec, xd[(i * 7 + 3) % POOL], flat, yd[(i * 7 + 3) % POOL], 1.0 / count, True, None, kb0))
    last = ((NK - 1) * 7 + 3) % POOL
    bwd_ms = replay_ms(lambda i: backbone_backward_raw(spec, xd[last], flat, saved, False, True, out=out, target=yd[last], gscale=2.0 / count,
                                                       gflat=gflat, bufs=kb1))
    clocks = sampler.stop()

    # ---- the serial floor: the same steps with every backbone forced to the plain serial kernels (tchunks = 1) — what the step
    # costs when no chunk boundary can be trusted (weights with a very long memory)
    saved_plans = []
    for mod in {id(m): m for m, *_ in trainer.chunk_calls()}.values():
        saved_plans.append((mod, getattr(mod, "time_chunks", None)))
        mod.time_chunks = (1, 1)
    trainer._graphs.clear()
    ms_serial, _ = run.timed(W, min(K, 200), flush)
    ms_serial /= min(K, 200)
    for mod, tc in saved_plans:
        if tc is None:
            del mod.time_chunks
        else:
            mod.time_chunks = tc
    trainer._graphs.clear()

    replicas_ok = None
    if world > 1:
        pa = torch.cat([p.detach().reshape(-1) for p in run.net.parameters()])
        allp = [torch.empty_like(pa) for _ in range(world)]
        dist.all_gather(allp, pa)
        replicas_ok = bool(all(torch.equal(allp[0], q) for q in allp))
    run.close()
    final_loss = float(losses[-1])
    run_multi = MULTI
    settle_steps = run.settle
    del run, xd, yd, xs_pin, ys_pin, saved, out, kb0, kb1
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs as extra keys: fixed GLOBAL batch split over the ranks (strong scaling), short runs
    secondary = {}
    if not args.no_secondary and args.workload == "c2a":
        for name in SECONDARY:
            w2 = WORKLOADS[name]
            if w2["B"] % world != 0 or (name == "c2a" and world == 1):
                continue
            try:
                r2 = Run(w2, dev, pg, world, rank, weak=False, settle=300)
                r2.settle_plan(max_passes=3)
                K2 = 30
                ms2, l2 = r2.timed(5, K2, flush)
                ci, _ = chunk_report(r2.trainer, r2.B, r2.T)
                secondary[name + ("_strong" if name == "c2a" else "")] = {
                    "workload": w2["name"], "dataset": w2["dataset"], "global_batch": r2.GB, "per_gpu_batch": r2.B, "frame_len": r2.T,
                    "scaling": "strong", "ms_per_step": ms2 / K2, "value": r2.GB * r2.T * K2 / (ms2 * 1e-3), "unit": "IQ samples/s",
                    "settle_steps": r2.settle, "timed_steps": K2, "pool_mb": r2.pool_mb, "final_loss": float(l2[-1]),
                    "time_chunks": [(c["cell"], c["dir"], c["chunks"], c["warmup_steps"], c["serial_reruns"]) for c in ci]}
                r2.close()
                del r2
                torch.cuda.empty_cache()
            except Exception as e:   # a secondary config must never take the headline down
                secondary[name] = {"error": repr(e)[:300]}
                if world > 1:
                    raise

    # ---- every other backbone of the reference + the layered path: whole train step, short runs (single-GPU lines only)
    if not args.no_secondary and args.workload == "c2a" and world == 1:
        others = {}
        for name in OTHER_BACKBONES:
            w2 = WORKLOADS[name]
            try:
                r2 = Run(w2, dev, pg, world, rank, weak=False, settle=40)
                K2 = 10
                ms2, l2 = r2.timed(3, K2, flush)
                others[name] = {"backbone": w2["name"], "ms_per_step": round(ms2 / K2, 4), "value": r2.GB * r2.T * K2 / (ms2 * 1e-3),
                                "final_loss": float(l2[-1])}
                r2.close()
                del r2
                torch.cuda.empty_cache()
            except Exception as e:
                others[name] = {"error": repr(e)[:300]}
        secondary["other_backbones"] = {"what": "whole train step (fwd + MSE + bwd + clip + AdamW) on real APA_200MHz frames, B=64 x T=2048 (qat_tres: the script's "
                                                "B=64 x T=200 cascade), 40 settle + 10 timed steps; these cells are not time-chunked (DESIGN.md 4.3/4.4)", "unit": "IQ samples/s", **others}

    if rank == 0:
        peaks, traffic = {}, {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        fam = {"dgru": "gru", "gru": "gru", "qgru": "gru", "deltagru": "delta", "deltagru_tcnskip": "delta"}.get(bb.cell, bb.cell)
        bwd_name = "odpd::gru_bwdf_kernel" if fam == "gru" else f"odpd::{fam}_bwd_kernel"      # GRU family: the lean fused backward (gru_family.cu)
        dom = (bwd_name, bwd_ms) if bwd_ms >= fwd_ms else (f"odpd::{fam}_fwd_kernel", fwd_ms)
        achieved = ALGO_BYTES_PER_SAMPLE_PER_KERNEL * B * T / (dom[1] * 1e-3) / 1e9
        plan_i = (0, 1) if "pa" not in wl else (1, 2)
        tr_entry = traffic.get(args.workload, {}).get(dom[0], {})
        line = {
            "metric": "IQ samples/sec/train-step", "value": world * B * T * K / (ms * 1e-3), "unit": "IQ samples/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": wl["dataset"],
            "gpu_launches": launches * K,
            "config": config_of(wl, world, True),
            "run": {"untimed_settle_steps": settle_steps, "untimed_settle": f"{settle_steps} real training steps over the seeded epoch permutation of the "
                    f"{wl['dataset']} stream (on-device framing) + {plan_settle} steps over the timed pool until the chunk controller made no change",
                    "steps_per_graph_replay": run_multi, "pool_batches": POOL, "pool_mb": POOL * 2 * B * T * 8 / 2**20, "cuda_graphs": bool(getattr(trainer, "use_graphs", False)),
                    "final_loss": final_loss, "loss_readback": "every timed step copies its loss device->host (pinned, async) inside the timed region; "
                    "value_sync_loss is the strict variant with loss.item() (host sync) after every step"},
            "value_sync_loss": {"value": world * B * T / (sync_ms * 1e-3), "ms_per_step": sync_ms},
            "serial_floor": {"ms_per_step": ms_serial, "value": world * B * T / (ms_serial * 1e-3),
                             "what": "same steps with every backbone on the plain serial kernels (tchunks=1): the cost when no time-chunk boundary holds"},
            "clocks": clocks,
            "e2e": {"value": world * B * T * K / e2e_s, "unit": "IQ samples/s", "ms_per_step": e2e_s / K * 1e3,
                    "h2d_bytes_per_step": 2 * B * T * 2 * 4, "d2h_bytes_per_step": 8,
                    "path": "NativeTrainStep.run_host_batches: per step pinned host (B,T,2) features+targets (real frames) -> cudaMemcpyAsync (side stream, "
                            "overlapping the previous block) -> fwd/bwd/optimizer kernels (blocks of run.steps_per_graph_replay steps per CUDA-graph "
                            "replay) -> async D2H of every step's loss, read one block later",
                    "ms_per_step_sequential": e2e_seq_ms},
            "e2e_indexed": {"value": world * B * T * K / e2e_idx_s, "unit": "IQ samples/s", "ms_per_step": e2e_idx_s / K * 1e3,
                            "h2d_bytes_per_step": 4 * B, "d2h_bytes_per_step": 8,
                            "path": "NativeTrainStep.step_indexed: the raw (N,2) training streams resident in HBM (as in a real epoch: 472 KB each, "
                                    "L2-resident by nature); per step the B frame start indices of the seeded permutation go pinned host -> device, "
                                    "the kernels read the stride-1 windows in place (OdpdDims.x_starts), the loss is read back one step later"},
            "kernels_per_step": (["<cell>_fwd_kernel (chunks)", "<cell>_fwd_kernel (verify)", "<cell>_bwd_kernel<DW> (chunks)", "<cell>_bwd_kernel<DW> (verify)",
                                  "reduce_partials_kernel (clip + AdamW in its last CTA at N=1; + dp_clip_adamw_kernel at N>1)"] if "pa" not in wl else
                                 ["dpd_fwd", "pa_fwd(+MSE)", "pa_bwd<dX>", "dpd_bwd<DW>", "(+1 verify launch per chunked call)",
                                  "reduce_partials_kernel (+ optimiser)", "(gmp: +1)"]),
            "time_chunks": chunk_info, "time_chunk_events": events_at_timed,
            "kernel_ms": {"fwd": fwd_ms, "bwd": bwd_ms},
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "traffic": tr_entry.get("dram_bytes_per_launch"), "traffic_source": tr_entry.get("source"),
                         "latency_view": {"timesteps": T,
                                          "chain_steps_per_cta_fwd": chunk_info[plan_i[0]]["steps_per_chunk"] + chunk_info[plan_i[0]]["warmup_steps"],
                                          "chain_steps_per_cta_bwd": chunk_info[plan_i[1]]["steps_per_chunk"] + chunk_info[plan_i[1]]["warmup_steps"],
                                          "ns_per_chain_step_fwd": fwd_ms * 1e6 / max(1, chunk_info[plan_i[0]]["steps_per_chunk"] + chunk_info[plan_i[0]]["warmup_steps"]),
                                          "ns_per_chain_step_bwd": bwd_ms * 1e6 / max(1, chunk_info[plan_i[1]]["steps_per_chunk"] + chunk_info[plan_i[1]]["warmup_steps"]),
                                          "note": "a T-step serial recurrence per sequence, cut into concurrently running chunks (DESIGN.md §4.1): bound by the latency of "
                                                  "one dependent step x the steps one CTA walks, not by bandwidth (SURVEY §8d)"}},
        }
        if dpc is not None:
            dpc["replicas_identical_after_run"] = replicas_ok
            line["dp_check"] = dpc
        if secondary:
            line["secondary"] = secondary
        if world == 1 and not args.no_cpu:
            try:
                line["gpu_reference"] = gpu_reference_leg(wl, B, dev)
                line["gpu_reference"]["native_over_gpu_reference"] = line["value"] / line["gpu_reference"]["value"]
                if "c3" in secondary and "error" not in secondary["c3"]:
                    g3 = gpu_reference_leg(WORKLOADS["c3"], 256, dev, n_warm=1, n_timed=2)
                    g3["native_over_gpu_reference"] = secondary["c3"]["value"] / g3["value"]
                    secondary["c3"]["gpu_reference"] = g3
            except Exception as e:
                line["gpu_reference_error"] = repr(e)[:300]
            r = cpu_port_leg(wl, B, seconds=args.cpu_seconds)
            leg = r.get("torch_port") or r["c_port"]
            line["cpu_baseline"] = {"value": leg["value"], "unit": "IQ samples/s", "cores": leg["cores"], "kind": "port", "sample": leg["sample"]}
            line["cpu_c_port"] = r["c_port"]
            if "torch_port_error" in r:
                line["torch_port_error"] = r["torch_port_error"]
            if "c3" in secondary and "error" not in secondary["c3"]:      # BASELINE configs[2]: the cascade's own CPU arm
                r3 = cpu_port_leg(WORKLOADS["c3"], 256, seconds=4.0).get("torch_port")
                if r3:
                    secondary["c3"]["cpu_baseline"] = {"value": r3["value"], "unit": "IQ samples/s", "cores": r3["cores"], "kind": "port",
                                                       "sample": r3["sample"], "native_over_cpu": secondary["c3"]["value"] / r3["value"]}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
