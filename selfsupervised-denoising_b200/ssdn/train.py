"""Training step, flat Adam optimiser, LR schedule and a compact DenoiserTrainer
(reference: ssdn/ssdn/train.py).

The hot loop of the reference is  zero_grad -> run_pipeline -> mean(loss).backward() -> Adam.step()
(train.py:197-202).  Here the parameters of all networks are views of one flat fp32 buffer, the engine
writes gradients into a matching flat buffer, and data parallelism is one process per GPU with exactly
one NCCL all-reduce of that buffer per step (the mean over ranks is folded into the Adam kernel)."""
from __future__ import annotations

import glob
import logging
import os
import re
from collections import defaultdict
from typing import Callable, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

import ssdn
from ssdn import _engine as E
from ssdn.datasets import FixedLengthSampler, HDF5Dataset, NoisyDataset, SamplingOrder, UnlabelledImageFolderDataset
from ssdn.denoiser import Denoiser
from ssdn.models import NoiseNetwork
from ssdn.params import ConfigValue, DatasetType, HistoryValue, PipelineOutput, StateValue
from ssdn.utils import Metric, MetricDict, TrackedTime, compute_ramped_lrate

DEFAULT_RUN_DIR = ssdn.cfg.DEFAULT_RUN_DIR
logger = logging.getLogger("ssdn.train")


class FlatAdam:
    """torch.optim.Adam(betas=(0.9, 0.99)) over the Denoiser's flat parameter buffer: one kernel per step."""

    def __init__(self, denoiser: Denoiser, lr: float = 1e-3, betas=(0.9, 0.99), eps: float = 1e-8):
        self.denoiser = denoiser
        self.param_groups = [{"lr": lr, "betas": tuple(betas), "eps": eps}]
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None

    def _ensure_state(self, flat):
        if self.exp_avg is None or self.exp_avg.shape != flat.shape or self.exp_avg.device != flat.device:
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.denoiser.parameters():
            p.grad = None

    def step(self, grad_scale: float = 1.0):
        flat = self.denoiser.flat_parameters()
        grads = self.denoiser.flat_gradients()
        self._ensure_state(flat)
        self.step_count += 1
        g = self.param_groups[0]
        E.adam_step(flat, grads, self.exp_avg, self.exp_avg_sq, g["lr"], self.step_count, g["betas"][0], g["betas"][1], g["eps"],
                    grad_scale, skip=self.denoiser.stale_flags())

    # ---- the same step with its scalars in device memory (what a CUDA graph of the step launches)
    def hyper_values(self, grad_scale: float = 1.0):
        """{lr / bias_correction1, beta1, beta2, eps, sqrt(bias_correction2), grad_scale} of the CURRENT step_count."""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        return [g["lr"] / (1.0 - b1 ** self.step_count), b1, b2, g["eps"], (1.0 - b2 ** self.step_count) ** 0.5, grad_scale]

    def step_dev(self, hyper: torch.Tensor):
        """Launch the update with scalars read from `hyper` (CUDA float[6]); the caller advances step_count and fills hyper."""
        flat = self.denoiser.flat_parameters()
        grads = self.denoiser.flat_gradients()
        self._ensure_state(flat)
        E.adam_step_dev(flat, grads, self.exp_avg, self.exp_avg_sq, hyper, skip=self.denoiser.stale_flags())

    def state_dict(self) -> Dict:
        """Wire format of ``torch.optim.Adam.state_dict()`` - what the reference stores under ``"optimizer"`` in a
        ``.training`` file (train.py:724): per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` keyed by the parameter's
        position in ``denoiser.parameters()``, and one parameter group.  The per-parameter moments are views of the two
        flat buffers, so the file holds each buffer once."""
        params = list(self.denoiser.parameters())
        state = {}
        if self.step_count > 0 and self.exp_avg is not None:
            off = 0
            for i, p in enumerate(params):
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[off:off + n].view(p.shape),
                            "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape)}
                off += n
        g = self.param_groups[0]
        group = {"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"], "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, state: Dict):
        """Accepts ``torch.optim.Adam`` state dictionaries (any torch version: integer or tensor ``step``) as written by
        the reference trainer or by state_dict() above, and the flat layout early versions of this package wrote."""
        dev = self.denoiser.device
        if "state" not in state:                                   # flat layout
            self.step_count = int(state["step"])
            self.param_groups = [dict(g) for g in state["param_groups"]]
            self.exp_avg = None if state["exp_avg"] is None else state["exp_avg"].to(dev)
            self.exp_avg_sq = None if state["exp_avg_sq"] is None else state["exp_avg_sq"].to(dev)
            return
        groups = state["param_groups"]
        if len(groups) != 1:
            raise ValueError("expected one Adam parameter group, found {}".format(len(groups)))
        g = groups[0]
        if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("the engine's Adam step has no weight decay / amsgrad / maximize")
        params = list(self.denoiser.parameters())
        if len(g["params"]) != len(params):
            raise ValueError("optimizer state holds {} parameters, the denoiser has {}".format(len(g["params"]), len(params)))
        self.param_groups = [{"lr": g["lr"], "betas": tuple(g["betas"]), "eps": g["eps"]}]
        per_param = state["state"]
        self.step_count = max((int(float(s["step"])) for s in per_param.values()), default=0)
        if not per_param:
            self.exp_avg = self.exp_avg_sq = None
            return
        total = sum(p.numel() for p in params)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for key, p in zip(g["params"], params):
            n = p.numel()
            s = per_param.get(key)
            if s is not None:
                if s["exp_avg"].numel() != n:
                    raise ValueError("optimizer state of parameter {} has {} elements, expected {}".format(key, s["exp_avg"].numel(), n))
                self.exp_avg[off:off + n].copy_(s["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(s["exp_avg_sq"].reshape(-1))
            off += n


def _backward_of_mean(loss: torch.Tensor, cache: Dict):
    """torch.mean(loss).backward() (train.py:199-200) without ATen's mean / fill / scale launches: the gradient of the mean is
    the constant 1 / numel, kept in a cached tensor and handed to backward() directly - four tiny kernels less per step, which
    is a percent of the step on the small per-GPU batches of strong scaling."""
    key = (tuple(loss.shape), loss.device, loss.dtype)
    g = cache.get(key)
    if g is None:
        g = cache[key] = torch.full(loss.shape, 1.0 / loss.numel(), device=loss.device, dtype=loss.dtype)
    loss.backward(g)


_MEAN_GRADS: Dict = {}


def train_step(denoiser: Denoiser, optimizer: FlatAdam, data: List, world_size: int = 1) -> Dict:
    """One optimisation step on this rank's shard of the batch.  With world_size > 1 every rank holds an equal shard,
    the local loss is the mean over the shard, and gradients are summed by ONE all-reduce then scaled by 1/world_size
    inside the Adam kernel - identical to the gradient of the mean over the global batch."""
    optimizer.zero_grad()
    denoiser.dp_world_size = world_size
    outputs = denoiser.run_pipeline(data)
    _backward_of_mean(outputs[PipelineOutput.LOSS], _MEAN_GRADS)
    if world_size > 1:
        dist.all_reduce(denoiser.flat_gradients_with_flags(), op=dist.ReduceOp.SUM)
    optimizer.step(grad_scale=1.0 / world_size)
    return outputs


class GraphedTrainStep:
    """train_step for a FIXED batch shape as ONE CUDA-graph launch: zero_grad -> run_pipeline -> mean(loss).backward() ->
    (gradient all-reduce) -> Adam -> loss to pinned host memory are captured once (after eager warm-up steps that also settle the
    operand scales) and replayed for every batch.  A step is ~110 kernel launches; on the small per-GPU batches of strong
    scaling their launch cost, not the device work, is what bounds the eager step.

    Input pipeline: the graph is captured TWICE, over two sets of static input buffers ("slots").  Call i stages its batch and
    the step's scalars (learning rate, Adam bias corrections, 1 / world_size: six floats in pinned memory) into slot i % 2 on a
    COPY stream and replays that slot's graph on the current stream once the copies have landed - so the host-to-device copy
    of batch i + 1 runs while step i computes, and nothing but graph launches (and two event operations) is ever queued on the
    compute stream: stream operations between two graph launches each cost a 10 - 40 us bubble (tests/dev_e2e_breakdown.py).
    Host batches should be pinned (torch's caching host allocator keeps a pinned block alive until the copy that reads it has
    run); device batches work too (the copy then waits for whatever produced them on the current stream).

    The returned outputs are the slot's static tensors: they stay valid until the call after the next one.  `loss_host()` is
    the step's per-sample loss in pinned host memory, written by the graph's last node (valid after a synchronize / the
    slot's `done` event).  Outputs of earlier EAGER steps must not be alive when the graph is captured (their autograd graph
    pins gradient accumulators to the default stream)."""

    N_SLOTS = 2

    def __init__(self, denoiser: Denoiser, optimizer: FlatAdam, example: List, world_size: int = 1, warmup: int = 3):
        self.denoiser, self.optimizer, self.world_size = denoiser, optimizer, world_size
        dev = denoiser.device
        to_dev = lambda t: t.to(dev).clone() if torch.is_tensor(t) and t.numel() else t     # noqa: E731
        md = example[NoisyDataset.METADATA] if len(example) > NoisyDataset.METADATA else {}

        def make_slot():
            return [to_dev(example[0]), to_dev(example[1]) if len(example) > 1 else None,
                    {k: to_dev(v) for k, v in md.items() if k != NoisyDataset.Metadata.CLEAN}]
        self.slots = [make_slot() for _ in range(self.N_SLOTS)]
        self.hyper = [torch.zeros(6, dtype=torch.float32, device=dev) for _ in range(self.N_SLOTS)]
        self.hyper_host = [torch.zeros(6, dtype=torch.float32).pin_memory() for _ in range(self.N_SLOTS)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(self.N_SLOTS)]     # the slot's inputs have landed (copy stream)
        self.done = [torch.cuda.Event() for _ in range(self.N_SLOTS)]      # the replay that read the slot has finished
        self.calls = 0
        self.last = 0
        self._mean_grads: Dict = {}          # allocated during the eager warm-up, i.e. before (not inside) the captures
        import gc
        gc.collect()
        cur = torch.cuda.current_stream(dev)
        self._stage(0, None)
        cur.wait_event(self.ready[0])
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for i in range(max(1, warmup)):            # eager steps: lazily created streams / attributes / settled operand scales
                if i:
                    self._stage(0, None)
                    side.wait_event(self.ready[0])
                out = self._body(0)
                loss_shape, loss_dtype = out[PipelineOutput.LOSS].shape, out[PipelineOutput.LOSS].dtype
                del out
                self.done[0].record(side)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graphs, self._outputs = [], []
        self._loss_host = [torch.empty(loss_shape, dtype=loss_dtype).pin_memory() for _ in range(self.N_SLOTS)]   # (not during capture)
        for s in range(self.N_SLOTS):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._body(s)
                # a memcpy node: the step's result reaches the host without a stream operation of its own
                self._loss_host[s].copy_(out[PipelineOutput.LOSS].detach(), non_blocking=True)
            self.graphs.append(g); self._outputs.append(out)
        # the captured launches themselves did not execute: run slot 0's once so that step_count and the weights agree
        self._stage(0, None)
        cur.wait_event(self.ready[0])
        self.graphs[0].replay()
        self.done[0].record(cur)
        self.calls = 1

    @property
    def outputs(self) -> Dict:
        """Outputs (static tensors) of the most recent step."""
        return self._outputs[self.last]

    @property
    def static(self) -> List:
        return self.slots[self.last]

    def loss_host(self) -> torch.Tensor:
        """Per-sample loss of the most recent step in pinned host memory (complete once that step has finished on the device)."""
        return self._loss_host[self.last]

    def _stage(self, s: int, data: Optional[List]):
        """Queue the copies of `data` (None: keep the slot's contents) and of the step's scalars into slot s on the copy stream."""
        self.optimizer.step_count += 1
        self.ready[s].synchronize()                     # the previous copy out of this slot's pinned scalars has run (two calls ago)
        self.hyper_host[s].copy_(torch.tensor(self.optimizer.hyper_values(1.0 / self.world_size), dtype=torch.float32))
        cur = torch.cuda.current_stream(self.denoiser.device)
        cs = self.copy_stream
        cs.wait_event(self.done[s])                     # the replay that last read this slot has finished
        slot = self.slots[s]
        with torch.cuda.stream(cs):
            if data is not None:
                md = data[NoisyDataset.METADATA] if len(data) > NoisyDataset.METADATA else {}
                pairs = [(slot[0], data[0])]
                if slot[1] is not None and torch.is_tensor(slot[1]) and slot[1].numel():
                    pairs.append((slot[1], data[1]))
                pairs += [(v, md[k]) for k, v in slot[2].items() if torch.is_tensor(v) and k in md]
                if any(src.is_cuda for _, src in pairs):
                    cs.wait_stream(cur)                 # device batches: whatever produced them on the current stream comes first
                for dst, src in pairs:
                    dst.copy_(src, non_blocking=True)
                    if src.is_cuda:
                        src.record_stream(cs)
            self.hyper[s].copy_(self.hyper_host[s], non_blocking=True)
            self.ready[s].record(cs)

    def _body(self, s: int):
        self.optimizer.zero_grad()
        self.denoiser.dp_world_size = self.world_size
        outputs = self.denoiser.run_pipeline(self.slots[s])
        _backward_of_mean(outputs[PipelineOutput.LOSS], self._mean_grads)
        if self.world_size > 1 and not os.environ.get("SSDN_DEV_SKIP_ALLREDUCE"):    # (developer timing switch: NOT a valid training step)
            dist.all_reduce(self.denoiser.flat_gradients_with_flags(), op=dist.ReduceOp.SUM)
        self.optimizer.step_dev(self.hyper[s])
        return outputs

    def __call__(self, data: List) -> Dict:
        s = self.calls % self.N_SLOTS
        self.calls += 1
        self._stage(s, data)
        cur = torch.cuda.current_stream(self.denoiser.device)
        cur.wait_event(self.ready[s])
        self.graphs[s].replay()
        self.done[s].record(cur)
        self.last = s
        return self._outputs[s]


def learning_rate(cfg: Dict, iteration: int) -> float:
    """The reference passes (RAMPDOWN, RAMPUP) into (ramp_up, ramp_down) (train.py:276-282) with the default values
    0.1 / 0.3 (cfg.py:18-19): the effective schedule ramps up over the first 10 % of the images and down over the last
    30 %.  Reproduced as is."""
    return compute_ramped_lrate(iteration, cfg[ConfigValue.TRAIN_ITERATIONS], cfg[ConfigValue.LR_RAMPDOWN_FRACTION],
                                cfg[ConfigValue.LR_RAMPUP_FRACTION], cfg[ConfigValue.LEARNING_RATE])


class RankShardSampler(torch.utils.data.Sampler):
    """Rank r of W reads positions cursor + r, cursor + r + W, ... of the wrapped sampler's global order.  The order itself is
    rank 0's (broadcast), so ranks need not share a seed; the trainer keeps the global cursor (images seen by all ranks)."""

    def __init__(self, base: FixedLengthSampler, rank: int, world: int):
        self.base, self.rank, self.world = base, rank, world

    def _order(self) -> SamplingOrder:
        order = iter(self.base)
        if self.world > 1 and dist.is_available() and dist.is_initialized():
            box = [order.state_dict() if self.rank == 0 else None]
            dist.broadcast_object_list(box, 0)
            if self.rank != 0:
                order = SamplingOrder.from_state_dict(box[0])
                self.base.for_next_iter(order)
        return order

    def __iter__(self):
        order = self._order()
        return iter(order.order[order.index + self.rank::self.world])

    def __len__(self) -> int:
        order = self.base.last_iter()
        start = order.index if order is not None else 0
        return max(0, (len(self.base) - start - self.rank + self.world - 1) // self.world)


class DenoiserTrainer:
    """Compact counterpart of the reference trainer: drives train_step over any iterable of NoisyDataset-style batches,
    keeps the iteration counter in IMAGES (train.py:221), accumulates the same metrics, snapshots and resumes.
    TensorBoard and image dumps of the reference are outside the hot path and are not reproduced."""

    def __init__(self, cfg: Dict, state: Optional[Dict] = None, runs_dir: str = DEFAULT_RUN_DIR, run_dir: str = None):
        self.runs_dir = os.path.abspath(runs_dir)
        self._run_dir = run_dir
        self.cfg = cfg
        if self.cfg:
            no_data = self.cfg.get(ConfigValue.TRAIN_DATA_PATH) is None and self.cfg.get(ConfigValue.TEST_DATA_PATH) is None
            ssdn.cfg.infer(self.cfg, model_only=no_data)
        self.state = state if state is not None else {}
        self._denoiser: Optional[Denoiser] = None
        self._optimizer: Optional[FlatAdam] = None
        self._train_iter: Optional[SamplingOrder] = None       # restored sample order waiting for a sampler (train.py:743,795-797)
        self.train_sampler: Optional[FixedLengthSampler] = None
        self.world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world_size > 1 else 0
        self.last_eval: Optional[Dict] = None

    @property
    def denoiser(self) -> Denoiser:
        return self._denoiser

    @denoiser.setter
    def denoiser(self, denoiser: Denoiser):
        self._denoiser = denoiser
        self.init_optimiser()

    def init_optimiser(self):
        self._optimizer = FlatAdam(self.denoiser, betas=[0.9, 0.99])

    def new_target(self, device: str = None):
        self.denoiser = Denoiser(self.cfg, device=device)
        self.init_state()
        self.sync_replicas()

    def sync_replicas(self):
        """Data-parallel runs (one process per GPU): every rank adopts rank 0's weights, Adam moments and step count, so that
        replicas start identical whatever each process seeded or loaded; train_step keeps them identical from then on (all
        ranks apply the same all-reduced gradient)."""
        if self.world_size <= 1 or self.denoiser is None or self.denoiser.device.type != "cuda":
            return
        flat = self.denoiser.flat_parameters()
        dist.broadcast(flat, 0)
        opt = self._optimizer
        meta = torch.tensor([opt.step_count, 0 if opt.exp_avg is None else 1], dtype=torch.int64, device=flat.device)
        dist.broadcast(meta, 0)
        opt.step_count = int(meta[0])
        if int(meta[1]):
            opt._ensure_state(flat)
            dist.broadcast(opt.exp_avg, 0)
            dist.broadcast(opt.exp_avg_sq, 0)

    def check_engine(self) -> int:
        """Device-side health of every live execution plan: raises EngineError if a kernel pipeline timed out (results since
        the last check are then garbage and must not be saved); returns the number of passes that ran with stale operand
        scales (their optimiser steps were skipped on the device)."""
        stale = 0
        if self.denoiser is None:
            return 0
        for net in self.denoiser._models.values():
            for plan in getattr(net, "_plans", {}).values():
                plan.check()
                stale += plan.scale_status()[2]
        if stale > getattr(self, "_stale_seen", 0):
            logger.warning("%d forward/backward passes ran with stale operand scales so far (their updates were skipped or repeated)", stale)
        self._stale_seen = stale
        return stale

    def init_state(self):
        self.state[StateValue.INITIALISED] = True
        self.state[StateValue.ITERATION] = 0
        # same containers as train.py:114-125 - they are pickled inside .training files and the reference indexes
        # TIMINGS with keys it has not created yet ("last_print", "eta")
        self.state[StateValue.HISTORY] = {HistoryValue.TRAIN: MetricDict(), HistoryValue.EVAL: MetricDict(),
                                          HistoryValue.TIMINGS: defaultdict(TrackedTime)}
        self.reset_metrics()

    def reset_metrics(self, eval: bool = True, train: bool = True):
        """Empty the metric accumulators; the sample counter "n" is a plain integer next to them (train.py:515-533)."""
        chosen = ([HistoryValue.TRAIN] if train else []) + ([HistoryValue.EVAL] if eval else [])
        for which in chosen:
            metrics = self.state[StateValue.HISTORY][which]
            metrics["n"] = 0
            for m in metrics.values():
                if isinstance(m, Metric):
                    m.reset()

    def attach_sampler(self, sampler: FixedLengthSampler):
        """Register the sampler that orders the training data, so that snapshots record its order and position and a
        resumed run continues the same order (train.py:721-723,795-797)."""
        self.train_sampler = sampler
        if self._train_iter is not None:
            sampler.for_next_iter(self._train_iter)
            self._train_iter = None

    @property
    def learning_rate(self) -> float:
        return learning_rate(self.cfg, self.state[StateValue.ITERATION])

    @property
    def optimizer(self) -> FlatAdam:
        for group in self._optimizer.param_groups:
            group["lr"] = self.learning_rate
        return self._optimizer

    def train(self, batches: Iterable = None, on_step: Callable[[int, Dict], None] = None, intervals: bool = None):
        """Consume batches until TRAIN_ITERATIONS images have been seen.  Without ``batches`` the training set named by the
        configuration is read through the reference's CPU loader (train_data); ``GpuNoisyPatches`` is the fast source.
        ``intervals`` (default: on when the trainer reads its own data) adds the reference's housekeeping at multiples of
        EVAL_INTERVAL / PRINT_INTERVAL / SNAPSHOT_INTERVAL images and the final snapshot + ``final-<config>.wt``
        (train.py:155-181, 223-231)."""
        if self.denoiser is None:
            self.new_target()
        if intervals is None:
            intervals = batches is None
        if batches is None:
            batches, _, _ = self.train_data()
        test_batches = None
        if intervals and self.cfg.get(ConfigValue.TEST_DATA_PATH):
            test_batches, _, _ = self.test_data()
        history = self.state[StateValue.HISTORY][HistoryValue.TRAIN]
        self.denoiser.train()
        for data in batches:
            if intervals:
                self._housekeeping(test_batches)
            if self.state[StateValue.ITERATION] >= self.cfg[ConfigValue.TRAIN_ITERATIONS]:
                break
            outputs = train_step(self.denoiser, self.optimizer, data, self.world_size)
            n = data[NoisyDataset.INPUT].shape[0]
            with torch.no_grad():
                history["n"] += n
                history["loss"] += outputs[PipelineOutput.LOSS]
                clean = self._clean(data)
                if clean is not None:
                    history["psnr_out"] += ssdn.utils.calculate_psnr(outputs[PipelineOutput.IMG_DENOISED], clean)
                    if PipelineOutput.IMG_MU in outputs:
                        history["psnr_mu_out"] += ssdn.utils.calculate_psnr(outputs[PipelineOutput.IMG_MU].contiguous(), clean)
                for key in (PipelineOutput.NOISE_STD_DEV, PipelineOutput.MODEL_STD_DEV):
                    if key in outputs:
                        history[key.value] += outputs[key] * 255
            self.state[StateValue.ITERATION] += n * self.world_size
            if on_step:
                on_step(self.state[StateValue.ITERATION], outputs)
        if intervals and self.state[StateValue.ITERATION] >= self.cfg[ConfigValue.TRAIN_ITERATIONS]:
            self._housekeeping(test_batches)
            if self.rank == 0:
                self.snapshot()
                self.snapshot(output_name="final-{}.wt".format(self.denoiser.config_name()), subdir="", model_only=True)

    def _housekeeping(self, test_batches: Iterable = None):
        """What the reference does at the top of every iteration (train.py:155-181): evaluate, report and reset the metric
        window, snapshot - each when the image counter is a multiple of its interval."""
        iteration = self.state[StateValue.ITERATION]
        history = self.state[StateValue.HISTORY]
        if getattr(self, "_housekept", None) == iteration:       # once per value of the image counter
            return
        self._housekept = iteration
        if test_batches is not None and iteration % self.cfg[ConfigValue.EVAL_INTERVAL] == 0:
            self.last_eval = self.evaluate(test_batches)
            self.denoiser.train()
        if iteration % self.cfg[ConfigValue.PRINT_INTERVAL] == 0:
            self.check_engine()
            history[HistoryValue.TIMINGS]["total"].update()
            if self.rank == 0:
                logger.info(self.state_str())
            self.reset_metrics()
        if iteration % self.cfg[ConfigValue.SNAPSHOT_INTERVAL] == 0 and self.rank == 0:
            self.snapshot()

    def state_str(self) -> str:
        """One line per metric window: image counter, learning rate, means of the accumulated training metrics and the last
        evaluation result."""
        train = self.state[StateValue.HISTORY][HistoryValue.TRAIN]
        parts = ["n={}".format(train["n"]), "lr={:.3e}".format(self.learning_rate)]
        for name, metric in train.items():
            if isinstance(metric, Metric) and not metric.empty():
                parts.append("{}={:.4f}".format(name, float(metric.accumulated().mean())))
        line = "[{:08d}] TRAIN | {}".format(self.state[StateValue.ITERATION], ", ".join(parts))
        if getattr(self, "last_eval", None):
            line += " | VALID " + ", ".join("{}={:.4f}".format(k, v) for k, v in self.last_eval.items())
        total = self.state[StateValue.HISTORY][HistoryValue.TIMINGS]["total"].total
        return line + " | " + (ssdn.utils.seconds_to_dhms(total) or "0s")

    def _clean(self, data):
        md = data[NoisyDataset.METADATA] if len(data) > NoisyDataset.METADATA else None
        if md and NoisyDataset.Metadata.CLEAN in md and md[NoisyDataset.Metadata.CLEAN] is not None:
            return md[NoisyDataset.Metadata.CLEAN].to(self.denoiser.device)
        return None

    def evaluate(self, batches: Iterable = None, output_callback: Callable[[int, Dict], None] = None) -> Dict:
        """Forward-only pass; returns mean PSNR of the denoised output (and of mu for SSDN) on the unpadded region.
        Without ``batches`` the test set named by the configuration is used (test_data)."""
        if batches is None:
            batches, _, _ = self.test_data()
        self.denoiser.eval()
        metrics = MetricDict()
        idx = 0
        with torch.no_grad():
            for data in batches:
                outputs = self.denoiser.run_pipeline(data)
                clean = self._clean(data)
                md = data[NoisyDataset.METADATA]
                for key, name in ((PipelineOutput.IMG_DENOISED, "psnr_out"), (PipelineOutput.IMG_MU, "psnr_mu_out")):
                    if key not in outputs or clean is None:
                        continue
                    imgs = NoisyDataset.unpad(outputs[key], md)
                    refs = NoisyDataset.unpad(clean, md)
                    vals = [ssdn.utils.calculate_psnr(i.contiguous()[None], r.contiguous()[None]) for i, r in zip(imgs, refs)]
                    metrics[name] += torch.cat(vals)
                if output_callback:
                    output_callback(idx, outputs)
                idx += data[NoisyDataset.INPUT].shape[0]
        self.denoiser.train()
        return {k: float(v.accumulated()) for k, v in metrics.items()}

    # ------------------------------------------------------------------ data sets named by the configuration (train.py:746-864)
    def _clean_images(self, path: str, kind: DatasetType, transform=None):
        if kind == DatasetType.FOLDER:
            return UnlabelledImageFolderDataset(path, channels=self.cfg[ConfigValue.IMAGE_CHANNELS], transform=transform, recursive=True)
        if kind == DatasetType.HDF5:
            return HDF5Dataset(path, transform=transform, channels=self.cfg[ConfigValue.IMAGE_CHANNELS])
        raise NotImplementedError("Dataset type not implemented")

    def _loader(self, dataset: NoisyDataset, sampler: FixedLengthSampler, batch_size: int, shard: bool = False):
        """shard: data-parallel training - this rank reads positions rank, rank + W, ... of the ONE global order (drawn by rank
        0) in batches of batch_size / W, so that the W ranks together consume exactly the reference's global mini-batches."""
        from torch.utils.data import DataLoader
        if shard and self.world_size > 1:
            if batch_size % self.world_size:
                raise ValueError("TRAIN_MINIBATCH_SIZE {} is not divisible by the {} data-parallel ranks".format(batch_size, self.world_size))
            sampler, batch_size = RankShardSampler(sampler, self.rank, self.world_size), batch_size // self.world_size
        return DataLoader(dataset, sampler=sampler, batch_size=batch_size, num_workers=self.cfg[ConfigValue.DATALOADER_WORKERS],
                          pin_memory=self.cfg[ConfigValue.PIN_DATA_MEMORY])

    def train_data(self):
        """(DataLoader, NoisyDataset, FixedLengthSampler) over random TRAIN_PATCH_SIZE crops of the training images, in a
        shuffled fixed-length order that snapshots record and resume."""
        from torchvision.transforms import RandomCrop
        cfg = self.cfg
        crop = RandomCrop(cfg[ConfigValue.TRAIN_PATCH_SIZE], pad_if_needed=True, padding_mode="reflect")
        images = self._clean_images(cfg[ConfigValue.TRAIN_DATA_PATH], cfg[ConfigValue.TRAIN_DATASET_TYPE], crop)
        dataset = NoisyDataset(images, cfg[ConfigValue.NOISE_STYLE], cfg[ConfigValue.ALGORITHM], pad_uniform=False,
                               pad_multiple=NoiseNetwork.input_wh_mul(), square=cfg[ConfigValue.BLINDSPOT], training_mode=True)
        _ = dataset[0]
        sampler = FixedLengthSampler(dataset, num_samples=cfg[ConfigValue.TRAIN_ITERATIONS], shuffled=True)
        self.attach_sampler(sampler)
        return self._loader(dataset, sampler, cfg[ConfigValue.TRAIN_MINIBATCH_SIZE], shard=True), dataset, sampler

    def test_data(self):
        """(DataLoader, NoisyDataset, FixedLengthSampler) over the whole test images, reflect-padded to one common size
        that is a multiple of 32 (square for blind-spot networks); test_length(name) images, cycling if the set is smaller."""
        cfg = self.cfg
        images = self._clean_images(cfg[ConfigValue.TEST_DATA_PATH], cfg[ConfigValue.TEST_DATASET_TYPE])
        dataset = NoisyDataset(images, cfg[ConfigValue.NOISE_STYLE], cfg[ConfigValue.ALGORITHM], pad_uniform=True,
                               pad_multiple=NoiseNetwork.input_wh_mul(), square=cfg[ConfigValue.BLINDSPOT], training_mode=False)
        _ = dataset[0]
        sampler = FixedLengthSampler(dataset, num_samples=ssdn.cfg.test_length(cfg[ConfigValue.TEST_DATASET_NAME]), shuffled=False)
        return self._loader(dataset, sampler, cfg[ConfigValue.TEST_MINIBATCH_SIZE]), dataset, sampler

    def set_train_data(self, path: str):
        self.cfg[ConfigValue.TRAIN_DATA_PATH] = path
        self.cfg[ConfigValue.TRAIN_DATASET_TYPE] = self.cfg[ConfigValue.TRAIN_DATASET_NAME] = None
        ssdn.cfg.infer_datasets(self.cfg)

    def set_test_data(self, path: str):
        self.cfg[ConfigValue.TEST_DATA_PATH] = path
        self.cfg[ConfigValue.TEST_DATASET_TYPE] = self.cfg[ConfigValue.TEST_DATASET_NAME] = None
        ssdn.cfg.infer_datasets(self.cfg)

    # ------------------------------------------------------------------ persistence (train.py:378-408, 711-745, 871-909)
    @property
    def run_dir_path(self) -> str:
        if self._run_dir is None:
            os.makedirs(self.runs_dir, exist_ok=True)
            ids = [int(m.group(1)) for d in os.listdir(self.runs_dir) if (m := re.match(r"^(\d+)-", d))]
            self._run_dir = "{:05d}-train-{}".format(max(ids) + 1 if ids else 0, self.denoiser.config_name())
        return os.path.join(self.runs_dir, self._run_dir)

    def state_dict(self) -> Dict:
        """The reference's ``.training`` layout (train.py:711-726): denoiser, trainer state, the sample order with its
        cursor set to the images actually processed, ``torch.optim.Adam`` state, CPU RNG state.  Without an attached
        sampler (the on-GPU input pipeline is a pure function of (seed, step) and needs none) the order is empty."""
        order = self.train_sampler.last_iter() if self.train_sampler is not None else self._train_iter
        order_state = dict(order.state_dict()) if order is not None else {"order": []}
        order_state["index"] = self.state[StateValue.ITERATION]
        return {"denoiser": self.denoiser.state_dict(), "state": self.state, "train_order_iter": order_state,
                "optimizer": self._optimizer.state_dict(), "rng": torch.get_rng_state()}

    def load_state_dict(self, state_dict, device: str = None):
        if isinstance(state_dict, str):
            state_dict = torch.load(state_dict, map_location="cpu", weights_only=False)
        self.denoiser = Denoiser.from_state_dict(state_dict["denoiser"], device=device)
        self.cfg = self.denoiser.cfg
        self.state = state_dict["state"]
        if "train_order_iter" in state_dict:
            self._train_iter = SamplingOrder.from_state_dict(state_dict["train_order_iter"])
            if self.train_sampler is not None:
                self.attach_sampler(self.train_sampler)
        self._optimizer.load_state_dict(state_dict["optimizer"])
        torch.set_rng_state(state_dict["rng"])
        self.sync_replicas()

    def snapshot(self, output_name: str = None, subdir: str = None, model_only: bool = False) -> str:
        """``<run>/training/model_<images>.training`` or, model_only, ``<run>/models/model_<images>.wt`` (train.py:378-408).
        Refuses to write (EngineError) when a device-side pipeline error was flagged since the last check."""
        self.check_engine()
        if subdir is None:
            subdir = "models" if model_only else "training"
        if output_name is None:
            output_name = "model_{:08d}.{}".format(self.state[StateValue.ITERATION], "wt" if model_only else "training")
        path = os.path.join(self.run_dir_path, subdir, output_name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save(self.denoiser.state_dict() if model_only else self.state_dict(), path)
        return path


def resume_run(run_dir: str, iteration: int = None, device: str = None) -> DenoiserTrainer:
    """Restore the trainer from the newest (or the requested) ``training/model_<iter>.training`` snapshot of a run."""
    snaps = {}
    for p in glob.glob(os.path.join(run_dir, "training", "*.training")):
        m = re.search(r"model_(\d+)\.training$", p)
        if m:
            snaps[int(m.group(1))] = p
    if not snaps:
        raise ValueError("Run directory contains no training files.")
    it = max(snaps) if iteration is None else iteration
    if it not in snaps:
        raise ValueError("Training file for iteration {} not found.".format(it))
    run_dir = os.path.abspath(run_dir)
    trainer = DenoiserTrainer(None, runs_dir=os.path.dirname(run_dir), run_dir=os.path.basename(run_dir))
    trainer.load_state_dict(snaps[it], device=device)
    for timing in trainer.state[StateValue.HISTORY][HistoryValue.TIMINGS].values():     # stale absolute times
        if isinstance(timing, TrackedTime):
            timing.forget()
    return trainer
