#!/usr/bin/env python
"""Headline benchmark: hyperspectral crops/s for one Hang2020 training step (forward + weighted
cross-entropy over the heads + backward [+ gradient all-reduce at N > 1]) on synthetic
(B, 369, 11, 11) crops -- BASELINE.json's metric on its config
"Full Hang2020 (spectral+spatial+3 heads), bands=369 classes=50, batch=1024 on 1xB200"
(1024 crops PER GPU at N > 1, i.e. the 8xB200 config's 8192 global batch: weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K --warmup W  # the reference's own module on host cores
    python bench.py --config cfg2|cfg5 ...                            # BASELINE configs 2 (spectral_network, B = 256) and
                                                                      # 5 (metadata_sensor_fusion, 32 sites, B = 512 per GPU)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field's definition.
`value`   : device-timed, crops resident in HBM.
`e2e`     : same step through the public nn.Module API with the crops in pinned HOST memory,
            H2D copy of every step's crops and D2H read of its loss inside the timed region.
`roofline`: dominant kernel (by CUDA-event time inside the timed region) against the measured
            peak of the pipe that bounds it, plus the whole step against the HBM roofline the
            metric names (184,521 algorithmic bytes per crop at B_local = 1024).
"""
from __future__ import annotations

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
# NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION / WARN; stdout carries exactly one JSON line here
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ.pop("NCCL_DEBUG")

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "hyperspectral crops/sec (fwd+bwd, 369-band 11x11)"
UNIT = "crops/s"
FILTERS = (32, 64, 128)


# ----------------------------------------------------------------------------- arithmetic
def param_count(bands: int, classes: int, branches=("spectral", "spatial"), sites: int = 0) -> int:
    """Trainable parameters (Hang2020(369, 50) = 731,836; spectral_network(369, 20) = 480,572;
    metadata_sensor_fusion(369, 32 sites, 50) = 738,280)."""
    n = 1 if len(branches) == 2 else 0          # alpha
    cin = bands
    for k, c in enumerate(FILTERS):
        conv = c * cin * 9 + c + 2 * c
        ks_spec = (3, 5, 7)[k]
        ks_spat = (7, 5, 3)[k]
        spec = 2 * (c * c * ks_spec + c) + classes * c + classes
        spat = (c + 1) + 2 * (ks_spat * ks_spat + 1) + classes * 4 * c + classes
        n += len(branches) * conv + (spec if "spectral" in branches else 0) + (spat if "spatial" in branches else 0)
        cin = c
    if sites:
        n += sites * 16 + 2 * 16 + classes * 16 + classes + classes * 2 * classes + classes
    return n


def algorithmic_bytes_per_crop(bands: int, classes: int, b_local: int, n_params: int, extra: int = 0) -> float:
    """crop read once + label + scores written once + params read and grads
    written once per step, amortised over the local batch (+ `extra` per-crop input bytes, e.g. the site id)."""
    return 4.0 * bands * 121 + 8 + 4 * classes + extra + 2.0 * 4 * n_params / b_local


def conv_flops(bands: int, nb: int = 2):
    """Nominal FLOPs per crop of each convolution GEMM (2*Cin*Cout*9*H*W, padded taps included),
    `nb` branches.  Keys match the library's stage names."""
    c1 = 2.0 * bands * 32 * nb * 9 * 121
    c2 = nb * 2.0 * 32 * 64 * 9 * 121
    c3 = nb * 2.0 * 64 * 128 * 9 * 25
    return {"fwd.conv1": c1, "fwd.conv2": c2, "fwd.conv3": c3,
            "bwd.conv1_wgrad": c1, "bwd.conv2_wgrad": c2, "bwd.conv3_wgrad": c3,
            "bwd.conv2_dgrad": c2, "bwd.conv3_dgrad": c3}


def step_flops_per_crop(bands: int, nb: int = 2) -> float:
    return sum(conv_flops(bands, nb).values())


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    # B200_PROFILING.md fallback
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- workload
def synth_batch(batch, bands, classes, seed):
    """crops U[0,1) (the loader's per-pixel min-max scaling, src/utils.py:49),
    labels uniform; per-rank seed."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, bands, 11, 11, generator=g)
    y = torch.randint(0, classes, (batch,), generator=g)
    return x, y


def loss_of(regime, joint, heads, y):
    """R2 (default, the north star): sum of CE over the six heads (three per branch).
    R1: the reference's TreeModel.training_step, CE(joint) (src/main.py:78)."""
    if regime == "R1":
        return F.cross_entropy(joint, y)
    return sum(F.cross_entropy(h, y) for h in heads)


# The three BASELINE configurations this file can time (configs[1] = cfg3 is the one the metric is quoted on).
CONFIGS = {
    "cfg3": dict(model="Hang2020", batch=1024, classes=50, sites=0, nb=2),
    "cfg2": dict(model="spectral_network", batch=256, classes=20, sites=0, nb=1),
    "cfg5": dict(model="metadata_sensor_fusion", batch=512, classes=50, sites=32, nb=2),
}


def workload_string(args, world):
    """One description of the workload for both arms (the reference arm runs a bounded per-step sample of it)."""
    c = CONFIGS[args.config]
    what = {"cfg3": f"Hang2020(bands={args.bands}, classes={args.classes}) fwd+CE({args.regime})+bwd",
            "cfg2": f"spectral_network(bands={args.bands}, classes={args.classes}) fwd+CE(sum of 3 heads)+bwd",
            "cfg5": f"metadata_sensor_fusion(bands={args.bands}, sites={c['sites']}, classes={args.classes}) fwd+CE+bwd, train()"}[args.config]
    return what + ("+grad all-reduce" if world > 1 else "") + f", {args.batch} crops per GPU per step"


def synth_sites(batch, sites, seed):
    g = torch.Generator().manual_seed(seed + 1)
    return torch.randint(0, sites, (batch,), generator=g)


def reference_module(args):
    """(module, kind): the reference's OWN model file (``kind = "reference"``: read from /root/reference in the build
    container, from the unmodified copy under oracle/_ref on the GPU box -- oracle/ref_loader.py), else the oracle port
    (``kind = "port"``: the same ATen CPU ops, pinned to the reference's outputs by tests/golden)."""
    from oracle import ref_loader
    bands, classes, c = args.bands, args.classes, CONFIGS[args.config]
    torch.manual_seed(0)
    if args.config == "cfg5":
        ref = ref_loader.load("metadata")
        if ref is not None:
            return ref.metadata_sensor_fusion(bands, c["sites"], classes).train(), "reference"
    else:
        ref = ref_loader.load("Hang2020")
        if ref is not None:
            cls = ref.Hang2020 if args.config == "cfg3" else ref.spectral_network
            return cls(bands, classes).train(), "reference"
    from oracle import hang2020_oracle as orc
    if args.config == "cfg5":
        from oracle import metadata_oracle as mo

        class _Port(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.table = mo.init_fusion_params(bands, c["sites"], classes, 0)
                self.params = torch.nn.ParameterList([torch.nn.Parameter(v) for k, v in self.table.items() if not mo.is_buffer(k)])
                it = iter(self.params)
                self.table = {k: (next(it) if not mo.is_buffer(k) else v) for k, v in self.table.items()}

            def forward(self, images, site):
                keep = torch.rand(images.shape[0], 16) >= 0.7
                return mo.fusion_forward(self.table, images, site, True, keep)
        return _Port().train(), "port"
    return orc.OracleModule("hang2020" if args.config == "cfg3" else "spectral", bands, classes, seed=0).train(), "port"


def reference_loss(args, m, kind, x, y, site):
    """The reference-side step's loss.  cfg3 R1: ``CE(Hang2020.forward(x))`` (src/main.py:77-78); cfg3 R2: CE summed over the
    six heads, obtained on the reference module the way SURVEY.md 0.3 prescribes (``m.spectral_network(x) +
    m.spatial_network(x)``, the two calls Hang2020.forward makes); cfg2: CE over the sub-network's three heads; cfg5: CE(out)."""
    if args.config == "cfg5":
        return F.cross_entropy(m(x, site), y)
    if args.config == "cfg2":
        return sum(F.cross_entropy(h, y) for h in m(x))
    if kind == "reference":
        if args.regime == "R1":
            return F.cross_entropy(m(x), y)
        return sum(F.cross_entropy(h, y) for h in m.spectral_network(x) + m.spatial_network(x))
    joint = m(x)
    return loss_of(args.regime, joint, m.heads, y)


def cpu_reference_rate(args, batch, steps, warmup, threads, model):
    """`steps` timed steps of `batch` crops of the reference's CPU implementation on `threads` host threads."""
    torch.set_num_threads(threads)
    m, kind = model
    x, y = synth_batch(batch, args.bands, args.classes, 0)
    site = synth_sites(batch, CONFIGS[args.config]["sites"], 0) if args.config == "cfg5" else None
    times = []
    for i in range(warmup + steps):
        for p in m.parameters():
            p.grad = None
        t0 = time.perf_counter()
        reference_loss(args, m, kind, x, y, site).backward()
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    total = sum(times)
    return batch * steps / total, total / steps


def reference_name(args, kind):
    if kind != "reference":
        return "oracle port"
    return "weecology/DeepTreeAttention src/models/" + ("metadata.py" if args.config == "cfg5" else "Hang2020.py") + " (unmodified)"


def run_reference(args):
    """--impl reference: the reference's own module on the box's host cores (kind "reference" when its three torch-only model
    files were copied to oracle/_ref by __graft_entry__.build(); kind "port" = the oracle restatement otherwise)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    batch = args.batch
    model = reference_module(args)
    kind = model[1]
    # bound the run: calibrate on a small sample, shrink the per-step sample until the whole
    # --steps/--warmup run fits in ~150 s
    rate0, _ = cpu_reference_rate(args, 64, 1, 1, threads, model)
    budget = 150.0
    while batch > 64 and (args.steps + args.warmup) * batch / rate0 > budget:
        batch //= 2
    rate, sec = cpu_reference_rate(args, batch, args.steps, args.warmup, threads, model)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args, args.gpus), "regime": args.regime, "batch_per_gpu": args.batch,
                   "sample_crops_per_step": batch, "where": "host cores"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{args.steps} steps x {batch} crops after {args.warmup} warm-up, {reference_name(args, kind)} on torch "
                                   f"{torch.__version__} CPU ATen (oneDNN), {threads} threads of {cores} cores"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_torch_eager(args):
    """--impl torch_eager: the incumbent GPU path -- the reference's layer stack (oracle port: the same ATen ops
    the reference's torch.nn modules dispatch, cuDNN / cuBLAS kernels) through stock PyTorch eager on cuda:0, with cuDNN TF32
    convolutions on (PyTorch's default) and off.  Prints {"tf32": crops/s, "fp32": crops/s}; run by the b200 arm in a
    subprocess so that nothing it does can disturb the headline measurement."""
    from oracle import hang2020_oracle as orc
    dev = torch.device(os.environ.get("DTA_BENCH_TORCH_DEVICE", "cuda:0"))
    m = orc.OracleModule("hang2020", args.bands, args.classes, seed=0).to(dev).train()
    x, y = synth_batch(args.batch, args.bands, args.classes, 0)
    x, y = x.to(dev), y.to(dev)
    out = {}
    for name, tf32 in (("tf32", True), ("fp32", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        times = []
        for i in range(args.warmup + args.steps):
            for p in m.parameters():
                p.grad = None
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            joint = m(x)
            loss_of(args.regime, joint, m.heads, y).backward()
            if dev.type == "cuda":
                torch.cuda.synchronize()
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        out[name] = args.batch * len(times) / sum(times)
    print(json.dumps(out), flush=True)


def torch_eager_gpu_baseline(args):
    """Runs run_torch_eager in a child process (bounded by a timeout) and returns its dict, or why it is unavailable."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "torch_eager", "--steps", "10", "--warmup", "3", "--batch", str(args.batch),
           "--bands", str(args.bands), "--classes", str(args.classes), "--regime", args.regime]
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            return {"unavailable": (res.stderr.strip().splitlines() or ["no output"])[-1][:200]}
        d = json.loads(lines[-1])
        return {"value_tf32": d["tf32"], "value_fp32": d["fp32"], "unit": UNIT,
                "how": f"extra: the reference's layer stack (oracle port, ATen/cuDNN kernels) through stock PyTorch {torch.__version__} eager "
                       "on the same GPU, 10 timed steps after 3 warm-up, cuDNN TF32 convolutions on (PyTorch default) / off"}
    except Exception as e:  # noqa: BLE001  (a baseline that cannot run must not take the benchmark down)
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


# ----------------------------------------------------------------------------- this repo's arm
class _FusionStep(torch.nn.Module):
    """cfg5 adapter: ``forward(images)`` of metadata_sensor_fusion with the step's site ids held in a device buffer (so the step
    has the one-tensor signature GraphedTrainStep captures; the e2e loop refreshes the buffer every step)."""

    def __init__(self, model, site):
        super().__init__()
        self.model, self.site = model, site

    def forward(self, images):
        return self.model(images, self.site)


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned staging buffers are
    allocated (first touch places their pages on that node): at N > 2 the end-to-end number was limited by crops crossing the
    inter-socket link on their way to PCIe.  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
        node = int(open(path + "/numa_node").read().strip())
        cpus = open(path + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        if ids:
            os.sched_setaffinity(0, ids)
        return {"numa_node": node, "cpus": cpus}
    except Exception as e:  # noqa: BLE001  (affinity is an optimisation, never a requirement)
        return {"numa_node": None, "why": f"{type(e).__name__}: {e}"[:120]}


def run_b200(args):
    import torch.distributed as dist
    from deeptreeattention_b200 import Hang2020 as H
    from deeptreeattention_b200 import _capi, distributed as D
    from deeptreeattention_b200.loss import cross_entropy_heads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    rank, world, local = D.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    affinity = bind_to_gpu_numa_node(local)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = CONFIGS[args.config]
    B, bands, classes, nb = args.batch, args.bands, args.classes, cfg["nb"]
    _capi.set_option(local, "overlap", args.overlap)
    _capi.set_option(local, "pdl", args.pdl)

    torch.manual_seed(0)                       # same replica on every rank (DDP semantics)
    x_host, y_host = synth_batch(B, bands, classes, seed=rank)
    site_host = synth_sites(B, cfg["sites"], rank) if cfg["sites"] else None
    if args.config == "cfg3":
        model = H.Hang2020(bands, classes).to(dev).train()
        core = model
        n_params = param_count(bands, classes)

        def heads_of(m, out):
            return [out] if args.regime == "R1" else m.head_scores
    elif args.config == "cfg2":
        model = H.spectral_network(bands, classes).to(dev).train()
        core = model
        n_params = param_count(bands, classes, ("spectral",))

        def heads_of(m, out):
            return out
    else:
        from deeptreeattention_b200 import metadata as M
        core = M.metadata_sensor_fusion(bands, cfg["sites"], classes).to(dev).train()
        model = _FusionStep(core, site_host.to(dev))
        n_params = param_count(bands, classes, sites=cfg["sites"])

        def heads_of(m, out):
            return [out]
    sync = D.GradSync(core)
    x_pin = [x_host.pin_memory(), x_host.clone().pin_memory()]
    y_pin = [y_host.pin_memory(), y_host.clone().pin_memory()]
    site_pin = [site_host.pin_memory(), site_host.clone().pin_memory()] if site_host is not None else None
    y_dev = y_host.to(dev)
    x_dev = x_host.to(dev)
    params = list(model.parameters())

    def loss_fn(m, out, y):
        # the library's fused weighted cross-entropy (== sum of F.cross_entropy over the heads, tests/test_gpu_parity.py)
        return cross_entropy_heads(heads_of(m, out), y)

    # --step fused (default where the loss is the sum over the network's own heads: cfg3 / R2, cfg2): the whole step is ONE
    # library call, train.fused_train_step -> dta_train_step (bit-identical to the three-call sequence below,
    # tests/test_train_step.py); --step autograd: model(x) -> cross_entropy_heads -> loss.backward() through torch.autograd.
    fused_ok = (args.config == "cfg3" and args.regime == "R2") or args.config == "cfg2"
    use_fused = args.step == "fused" and fused_ok
    if use_fused:
        from deeptreeattention_b200.train import GraphedFusedTrainStep, fused_train_step

        def train_step(xd, yd=None):
            loss = fused_train_step(model, xd, y_dev if yd is None else yd)
            sync.sync()
            return loss
    else:
        def train_step(xd, yd=None):
            for p in params:
                p.grad = None
            loss = loss_fn(model, model(xd), y_dev if yd is None else yd)
            loss.backward()
            sync.sync()
            return loss

    step_fn = train_step
    graphed = None
    if args.graph:
        from deeptreeattention_b200.graph import GraphedTrainStep
        if use_fused:
            graphed = GraphedFusedTrainStep(model, x_dev, y_dev, after_backward=sync.sync)
        else:
            graphed = GraphedTrainStep(model, x_dev, y_dev, loss_fn, after_backward=sync.sync)

        def step_fn(xd, yd=None):
            return graphed(xd, yd)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # resident input of the device-timed legs: the captured graph's own input buffer (no copy inside the timed region)
    x_res = graphed.x if graphed is not None else x_dev
    crops_bytes = B * bands * 484
    flush = None
    if crops_bytes <= 126e6:            # crops fit in L2: flush it between timed steps (a write larger than L2), per-step events
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def timed_replays(n):
        """n steps on resident crops, CUDA events on the launching stream; ms summed over the steps, max over ranks.  With an
        L2 flush every step is bracketed on its own and the flush sits outside the brackets."""
        barrier(); torch.cuda.synchronize()
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                step_fn(x_res)
            e1.record()
            torch.cuda.synchronize(); barrier()
            return max_over_ranks(e0.elapsed_time(e1))
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            flush.zero_()
            a.record()
            step_fn(x_res)
            b.record()
        torch.cuda.synchronize(); barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_fn(x_res)
    torch.cuda.synchronize()
    # kernels launched per step by the library: difference of the context's cumulative launch counter around one eager step
    l0 = _capi.get_option(local, "launches_total")
    train_step(x_dev)
    torch.cuda.synchronize()
    launches = (_capi.get_option(local, "launches_total") - l0) * args.steps
    if not args.graph:
        _capi.set_option(local, "profile", 1)
        _capi.profile_read(local, reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = timed_replays(args.steps)
    clock_rec = clocks.stop() if rank == 0 else None
    if args.graph:
        # per-stage CUDA-event timing cannot run inside a replayed graph: same kernels, same buffers, one
        # eager pass of `steps` steps right after the timed region, events on the launching stream
        _capi.set_option(local, "profile", 1)
        _capi.profile_read(local, reset=True)
        for _ in range(args.steps):
            train_step(x_dev)
        torch.cuda.synchronize()
    stages = _capi.profile_read(local, reset=True)
    _capi.set_option(local, "profile", 0)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- sustained: the same replays for >= args.sustain seconds, clocks and power sampled throughout -----------------
    sustained = None
    if args.sustain > 0:
        n_sus = max(args.steps, int(args.sustain * 1e3 / ms_step) + 1)
        sclk = ClockSampler(local)
        if rank == 0:
            sclk.start()
        ms_sus = timed_replays(n_sus)
        srec = sclk.stop() if rank == 0 else None
        sustained = {"value": world * B * n_sus / (ms_sus * 1e-3), "unit": UNIT, "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "clocks": srec}

    # ---- end to end through the public API, crops AND labels in pinned host memory ------------------------------------
    copy_stream = torch.cuda.Stream(dev)
    x_buf = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    y_buf = [torch.empty_like(y_dev), torch.empty_like(y_dev)]
    site_buf = [torch.empty_like(model.site), torch.empty_like(model.site)] if site_pin is not None else None
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    h2d_bytes = x_dev.numel() * 4 + y_dev.numel() * 8 + (model.site.numel() * 8 if site_pin is not None else 0)

    def prefetch(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            x_buf[s].copy_(x_pin[s], non_blocking=True)
            y_buf[s].copy_(y_pin[s], non_blocking=True)
            if site_pin is not None:
                site_buf[s].copy_(site_pin[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        for s in (0, 1):
            consumed[s].record(torch.cuda.current_stream())
        prefetch(0)
        last = 0.0
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)                      # next step's crops cross PCIe under this step's math
            torch.cuda.current_stream().wait_event(ready[i & 1])
            if site_pin is not None:
                model.site.copy_(site_buf[i & 1], non_blocking=True)
            loss = step_fn(x_buf[i & 1], y_buf[i & 1])
            consumed[i & 1].record(torch.cuda.current_stream())
            last = loss.item()                       # D2H read of the step's result
        return last

    e2e_loop(3)
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * args.steps / t_e2e

    def finish():
        """Leave without tearing NCCL down: destroying the communicator while captured graphs that hold NCCL kernels are
        alive can block forever; every rank has finished its work by the barrier, so a hard exit is safe."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    # ---- same, but the crops cross PCIe as raw int16 and are preprocessed on the device -------------
    from deeptreeattention_b200.data import preprocess_crops
    g = torch.Generator().manual_seed(100 + rank)
    raw_pin = [torch.randint(0, 10000, (B, bands, 11, 11), generator=g, dtype=torch.int16).pin_memory() for _ in range(2)]
    raw_buf = [torch.empty((B, bands, 11, 11), dtype=torch.int16, device=dev) for _ in range(2)]
    x_stage = graphed.x if args.graph else torch.empty_like(x_dev)

    def prefetch_raw(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            raw_buf[s].copy_(raw_pin[s], non_blocking=True)
            y_buf[s].copy_(y_pin[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_raw_loop(n):
        for s in (0, 1):
            consumed[s].record(torch.cuda.current_stream())
        prefetch_raw(0)
        last = 0.0
        for i in range(n):
            if i + 1 < n:
                prefetch_raw(i + 1)
            torch.cuda.current_stream().wait_event(ready[i & 1])
            preprocess_crops(raw_buf[i & 1], clip=0, out=x_stage)   # the bench model keeps all 369 bands: no clipping
            loss = step_fn(x_stage, y_buf[i & 1])
            consumed[i & 1].record(torch.cuda.current_stream())
            last = loss.item()
        return last

    e2e_raw_loop(3)
    torch.cuda.synchronize(); barrier()
    t0 = time.perf_counter()
    e2e_raw_loop(args.steps)
    torch.cuda.synchronize()
    t_raw = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_raw_value = world * B * args.steps / t_raw

    # ---- optimizer beside the path (extra): the same step with one fused Adam launch over all parameters captured behind the
    #      gradient exchange (src/main.py:135-136), device-timed like `value` ------------------------------------------------
    from deeptreeattention_b200.optim import FusedAdam
    adam_ms = None
    if args.graph and world == 1 and args.config == "cfg3":       # N = 1 only
        opt = FusedAdam(model.parameters(), lr=1e-4, capturable=True)
        if use_fused:
            graphed_opt = GraphedFusedTrainStep(model, x_dev, y_dev, after_backward=sync.sync, optimizer=opt)
        else:
            graphed_opt = GraphedTrainStep(model, x_dev, y_dev, loss_fn, after_backward=sync.sync, optimizer=opt)
        for _ in range(3):
            graphed_opt(x_dev)
        barrier(); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            graphed_opt(x_dev)
        a1.record()
        torch.cuda.synchronize(); barrier()
        adam_ms = max_over_ranks(a0.elapsed_time(a1)) / args.steps

    # ---- N > 1: is the gradient exchange the driver just timed RIGHT?  One eager backward without the exchange, then the
    #      path the timed steps used (peer kernel) against NCCL's AVG all-reduce and against the float64 mean of the gathered
    #      per-rank gradients; rank-identical bits via MAX == MIN over ranks ---------------------------------------------
    sync_check = None
    if world > 1:
        def flat_of():
            fl = torch.cat([p.grad.detach().reshape(-1).float() for p in params if p.grad is not None and p.dtype == torch.float32])
            db = [p.grad.detach().reshape(-1).double() for p in params if p.grad is not None and p.dtype == torch.float64]
            return torch.cat([fl.double()] + db).clone()

        # the path the timed steps used (exchange inside the backward pass when registered, else one kernel after it)
        for p in params:
            p.grad = None
        loss_fn(model, model(x_dev), y_dev).backward()
        sync.sync()
        torch.cuda.synchronize()
        mine = flat_of()
        path = sync.last_path
        # the same step's LOCAL gradients: exchange switched off (a fresh GradSync without peer buffers would reduce them with
        # NCCL; here nothing is reduced), then NCCL's AVG and the float64 mean of the gathered per-rank gradients
        peer = sync.peer
        core.__dict__.pop("_grad_buffers", None)
        for p in params:
            p.grad = None
        loss_fn(model, model(x_dev), y_dev).backward()
        torch.cuda.synchronize()
        local_flat = flat_of()
        if peer is not None:
            core.__dict__["_grad_buffers"] = (peer.flat, peer.alpha)
        nccl32 = local_flat.float()
        dist.all_reduce(nccl32, op=dist.ReduceOp.AVG)
        gathered = [torch.empty_like(local_flat) for _ in range(world)]
        dist.all_gather(gathered, local_flat)
        mean64 = torch.stack(gathered).mean(0)
        hi, lo = mine.clone(), mine.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        scale = float(mean64.abs().max())
        sync_check = {"path": path, "world": world, "elements": int(mine.numel()),
                      "max_abs_vs_nccl_avg": float((mine - nccl32.double()).abs().max()), "max_rel": float((mine - nccl32.double()).abs().max()) / scale,
                      "max_rel_vs_fp64_mean": float((mine - mean64).abs().max()) / scale,
                      "nccl_max_rel_vs_fp64_mean": float((nccl32.double() - mean64).abs().max()) / scale,
                      "rank_identical_bits": bool(torch.equal(hi, lo)), "grad_scale": scale,
                      "note": "float32 gradients + alpha's float64 gradient; max_rel = |exchange - NCCL AVG(float32)| / max|mean|"}

    if rank != 0:
        finish()
        return

    # ---- roofline ---------------------------------------------------------------------------
    peaks = measured_peaks()
    flops = conv_flops(bands, nb)
    stage_ms = {k: v[0] / args.steps for k, v in stages.items()}             # ms per step (all launches of the stage)
    launch_ms = {k: v[0] / max(v[1], 1) for k, v in stages.items()}          # ms per launch
    dominant = max(stage_ms, key=stage_ms.get) if stage_ms else None
    bytes_per_crop = algorithmic_bytes_per_crop(bands, classes, B, n_params, 8 if cfg["sites"] else 0)
    hbm_roof = peaks["hbm_gbs"] * 1e9 / bytes_per_crop
    # the stages are timed in a short eager pass at burst clocks: burst peak; the sustained run below is held to the sustained one
    roof = {"bound": "tensor", "kernel": dominant, "unit": "TFLOP/s", "peak": peaks["bf16_tflops"],
            "peak_source": f"{peaks['source']} bf16 dense, burst (stages timed in a {args.steps}-step pass)", "traffic": None}
    if dominant is not None:
        d_ms = launch_ms[dominant]
        d_flops = flops.get(dominant, 0.0) * B
        roof["launch_ms"] = d_ms
        roof["share_of_step"] = stage_ms[dominant] / ms_step
        roof["achieved"] = d_flops / (d_ms * 1e-3) / 1e12 if d_ms > 0 else None
        roof["frac"] = roof["achieved"] / roof["peak"] if roof["achieved"] else None
        roof["frac_of_sustained_peak"] = roof["achieved"] / peaks["bf16_tflops_sustained"] if roof["achieved"] else None
        roof["flops_per_launch"] = d_flops
    step_fl = step_flops_per_crop(bands, nb)
    roof["step_hbm"] = {"bound": "hbm", "achieved": (value / world) * bytes_per_crop / 1e9, "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": (value / world) / hbm_roof, "bytes_per_crop": bytes_per_crop,
                        "roofline_crops_per_s_per_gpu": hbm_roof}
    roof["step_tensor"] = {"achieved": (value / world) * step_fl / 1e12, "peak": peaks["bf16_tflops"],
                           "unit": "TFLOP/s", "frac": (value / world) * step_fl / 1e12 / peaks["bf16_tflops"],
                           "flops_per_crop": step_fl}
    if sustained is not None:
        sustained["step_tensor_frac_of_sustained_peak"] = (sustained["value"] / world) * step_fl / 1e12 / peaks["bf16_tflops_sustained"]
        sustained["step_hbm_frac"] = (sustained["value"] / world) / hbm_roof
    roof["stage_timing"] = ("eager pass after the timed region (graph replay cannot carry events)" if args.graph
                            else "inside the timed region")
    roof["stages_ms_per_step"] = {k: round(v, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1])}
    traffic_file = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            t = json.load(f)
        if dominant in t.get("kernels", {}):
            roof["traffic"] = t["kernels"][dominant].get("dram_bytes_per_launch")
        elif t.get("kernel") == dominant:
            roof["traffic"] = t.get("dram_bytes_per_launch")

    # ---- CPU baseline (bounded sample, rank 0, N = 1 only) ------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ref_model = reference_module(args)
        rate0, _ = cpu_reference_rate(args, 64, 1, 1, cores, ref_model)
        cb = B
        while cb > 64 and 3 * cb / rate0 > 25.0:
            cb //= 2
        rate, _ = cpu_reference_rate(args, cb, 2, 1, cores, ref_model)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": ref_model[1],
               "sample": f"2 timed steps x {cb} crops (1 warm-up) of the same workload, {reference_name(args, ref_model[1])}"
                         f" on torch {torch.__version__} CPU ATen, {cores} threads"}

    torch_gpu = torch_eager_gpu_baseline(args) if (world == 1 and not args.no_cpu and args.config == "cfg3") else None

    pcie_gbs = e2e_value / world * h2d_bytes / B / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_string(args, world), "name": args.config,
                   "regime": args.regime, "launch": "cuda-graph replay" if args.graph else "eager",
                   "step": "one dta_train_step call (train.fused_train_step)" if use_fused else "model(x) -> cross_entropy_heads -> backward() through autograd",
                   "batch_per_gpu": B, "global_batch": B * world,
                   "side_stream_overlap": int(args.overlap), "programmatic_dependent_launch": bool(args.pdl), "parallelism": f"dp{world}", "gradient_exchange": sync.last_path if world > 1 else None, "l2": f"crops per step = {crops_bytes / 1e6:.0f} MB > 126 MB L2 (no flush needed)" if flush is None
                   else f"crops per step = {crops_bytes / 1e6:.0f} MB < 126 MB L2: a 256 MB write flushes L2 before every timed step (per-step events, flush not timed)",
                   "host_affinity": affinity},
        "roofline": roof, "cpu_baseline": cpu, "torch_eager_gpu_baseline": torch_gpu, "clocks": clock_rec,
        "sustained": sustained,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "pcie_gbs_per_gpu": pcie_gbs,
                "how": "pinned host crops + labels" + (" + site ids" if site_pin is not None else "") +
                       " -> double-buffered H2D on a copy stream -> the same training step -> loss.item(); fp32 crops make this "
                       "PCIe-bound: pcie_gbs_per_gpu is the achieved host->device rate (PCIe 5 x16 delivers ~55-57 of its 64 GB/s)"},
        "e2e_raw_int16": {"value": e2e_raw_value, "unit": UNIT, "h2d_bytes_per_step": raw_buf[0].numel() * 2 + y_dev.numel() * 8, "d2h_bytes_per_step": 4,
                          "how": "second e2e figure (tests/test_preprocess.py: bit-identical to the sklearn path): raw int16 crops + labels from "
                                 "pinned host memory -> H2D -> on-device preprocess_crops (per-pixel min-max, src/utils.py:36-57) -> same step"},
        "with_adam": None if adam_ms is None else {
            "value": world * B / (adam_ms * 1e-3), "unit": UNIT, "ms_per_step": adam_ms,
            "how": "extra: the same step plus one FusedAdam launch over all parameters (src/main.py:135-136) captured in the graph"},
        "grad_sync_check": sync_check,
        "gpu_launches": launches,
    }
    print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_eager"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS), help="BASELINE.json configuration (default cfg3: the headline)")
    ap.add_argument("--batch", type=int, default=None, help="crops per GPU per step (default: the configuration's)")
    ap.add_argument("--bands", type=int, default=369)
    ap.add_argument("--classes", type=int, default=None)
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back replays for the `sustained` key (0 = skip)")
    ap.add_argument("--regime", default="R2", choices=["R1", "R2"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--graph", type=int, default=1, help="1: replay the step as one CUDA graph (default), 0: eager launches")
    ap.add_argument("--pdl", type=int, default=1, help="library option \"pdl\": 1 = programmatic dependent launch between kernels (default)")
    ap.add_argument("--step", default="fused", choices=["fused", "autograd"],
                    help="fused: one dta_train_step call per step where the loss is the sum over the network's heads (default); "
                         "autograd: model(x), cross_entropy_heads, loss.backward() through torch.autograd")
    ap.add_argument("--overlap", type=int, default=2,
                    help="library option \"overlap\": 0 = caller's stream only, 1 = one side stream, 2 = + auxiliary stream (default)")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = CONFIGS[args.config]["batch"]
    if args.classes is None:
        args.classes = CONFIGS[args.config]["classes"]
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_eager":
        run_torch_eager(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
