"""The SSOD training step with the reference's flow (trainer/ssod_trainer.py:587-680 train_instance,
:458-488 update_optimizer, trainer/trainer.py:193-251 build_optimizer) for `model_type == 'yolov5'`:

  teacher-EMA forward (native tcgen05 engine) -> NMS + pseudo labels (native, device resident) ->
  student forward on cat(labeled, strong-aug unlabeled) -> ComputeLoss + ComputeStudentMatchLoss (native fused
  fwd/bwd) -> backward -> [NCCL all-reduce of the student gradients] -> SGD-Nesterov -> ema / semi-ema update
  (native fused 5-stream kernel).

Data parallelism: one process per GPU, per-rank batch and per-rank BN statistics exactly like the
reference's DDP without SyncBN; the only collective is one SUM all-reduce of the flattened gradient arena per step
(loss*WORLD_SIZE followed by DDP's mean == sum of per-rank gradients); teachers stay bit-identical on all ranks
because the reduced gradients are.
"""
import math

import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .domain_loss import DomainLoss, TargetLoss
from .ema import ModelEMA, CosineEMA, SemiSupModelEMA, update_ema_pair, next_pair_decays, ema_scalars
from .loss import ComputeLoss
from .model import Model, SupModel
from .optim import FusedSGD
from .parallel import GradArena
from .pseudo_label import FairPseudoLabel
from .ssod_loss import ComputeStudentMatchLoss


def one_cycle(y1=0.0, y2=1.0, steps=100):  # reference utils/general.py:480-482
    return lambda x: ((1 - math.cos(x * math.pi / steps)) / 2) * (y2 - y1) + y1


class SSODTrainerStep:
    def __init__(self, cfg, device, rank=-1, world_size=1, epochs=None, batch_size=None, amp_dtype=torch.bfloat16,
                 pseudo_label_stats=None, nb=None, start_epoch=0):
        """pseudo_label_stats (LabelMatch only): dict(target_data_len, label_num_per_image, cls_ratio_gt) that the reference
        derives from its datasets (ssod_trainer.py:71).  nb = batches per epoch (len(train_loader)); it only sizes the
        warm-up window exactly like trainer/trainer.py:372-376 (nb=None: the 1000-iteration floor)."""
        self.cfg, self.device = cfg, device
        self.RANK, self.WORLD_SIZE = rank, world_size
        self.epochs = epochs if epochs is not None else cfg.epochs
        self.epoch = 0
        self.batch_size = batch_size if batch_size is not None else cfg.Dataset.batch_size
        self.amp_dtype = amp_dtype
        self.model = Model(cfg).to(device)
        self.model_type = self.model.model_type
        self.ema = ModelEMA(self.model)
        if cfg.hyp.burn_epochs > 0:
            self.semi_ema = None
        elif cfg.SSOD.cosine_ema:
            self.semi_ema = CosineEMA(self.ema.ema, decay_start=cfg.SSOD.ema_rate, total_epoch=self.epochs)
        else:
            self.semi_ema = SemiSupModelEMA(self.ema.ema, cfg.SSOD.ema_rate)
        self.fixed_accumulate = cfg.SSOD.fixed_accumulate
        self.build_optimizer(cfg)
        self.compute_loss = ComputeLoss(self.model, cfg)
        self.compute_un_sup_loss = ComputeStudentMatchLoss(self.model, cfg)
        self.domain_loss, self.target_loss = DomainLoss(), TargetLoss()       # ssod_trainer.py:262-263
        if getattr(cfg.SSOD, "pseudo_label_type", "FairPseudoLabel") == "LabelMatch":      # ssod_trainer.py:69-71
            from .labelmatch import LabelMatch
            ps = pseudo_label_stats or {}
            nc = cfg.Dataset.nc
            self.pseudo_label_creator = LabelMatch(cfg, int(ps.get("target_data_len", 0) / max(world_size, 1)), ps.get("label_num_per_image", 7.0),
                                                   ps.get("cls_ratio_gt", np.full(nc, 1.0 / nc)))
        else:
            self.pseudo_label_creator = FairPseudoLabel(cfg)
        self.da_loss_weights = cfg.SSOD.da_loss_weights
        self.last_opt_step = -1
        # trainer/trainer.py:372-376: number of warm-up iterations = max(warmup_epochs * nb, 1000), capped at half the run
        self.nb = nb
        if cfg.hyp.warmup_epochs > 0:
            self.nw = max(round(cfg.hyp.warmup_epochs * (nb or 0)), 1000)
            if nb:
                self.nw = min(self.nw, (self.epochs - start_epoch) / 2 * nb)
        else:
            self.nw = -1
        self._arena = None
        self.last = {}
        self.profile = False     # record CUDA events at the phase boundaries of train_instance
        self.phase_events = []
        self._graph = None       # captured CUDA graph of the whole step (train_instance_graphed)
        self._ema_scalars_dev = None
        self._teacher_stream = None
        self._teacher_keep = None
        self._bn_sync = None

    # trainer/trainer.py:193-217
    def build_optimizer(self, cfg):
        nbs = 64
        self.accumulate = max(round(nbs / self.batch_size), 1)
        weight_decay = cfg.hyp.weight_decay * self.batch_size * self.accumulate / nbs
        g_bnw, g_w, g_b = [], [], []
        for v in self.model.modules():
            if hasattr(v, 'bias') and isinstance(v.bias, nn.Parameter):
                g_b.append(v.bias)
            if isinstance(v, nn.BatchNorm2d):
                g_bnw.append(v.weight)
            elif hasattr(v, 'weight') and isinstance(v.weight, nn.Parameter):
                g_w.append(v.weight)
        self.optimizer = FusedSGD(g_b, lr=cfg.hyp.lr0, momentum=cfg.hyp.momentum, nesterov=True)   # one launch, zeroes the grads
        self.optimizer.add_param_group({'params': g_w, 'weight_decay': weight_decay})
        self.optimizer.add_param_group({'params': g_bnw})
        if cfg.linear_lr:
            self.lf = lambda x: (1 - x / (self.epochs - 1)) * (1.0 - cfg.hyp.lrf) + cfg.hyp.lrf
        else:
            self.lf = one_cycle(1, cfg.hyp.lrf, self.epochs)
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lr_lambda=self.lf)
        self.warmup_bias_lr, self.warmup_momentum, self.momentum = cfg.hyp.warmup_bias_lr, cfg.hyp.warmup_momentum, cfg.hyp.momentum

    # ---- gradient arena: all student gradients live in one flat fp32 buffer -> ONE all-reduce per step ----
    def _ensure_arena(self):
        if self._arena is None:
            bb = self.model.backbone
            # backward-completion order + chunk boundaries at the two autograd marks of YoloV5BackBone.forward:
            # [netD, head, neck, sppf, stage5_2] | [stage5_1, stage4_2] | [stage4_1 ... stem]
            self._arena = GradArena(self.model.parameters(), self.device, reverse=True,
                                    chunk_ends=[bb.stage5_2.cv1.conv.weight, bb.stage4_2.cv1.conv.weight])
            if self.WORLD_SIZE > 1:
                from . import autograd_conv as ac
                side = lambda: ac.WGRAD_SIDE["stream"] if ac.WGRAD_SIDE["dirty"] else None  # noqa: E731
                bb.grad_marks = tuple((lambda k=k: self._arena.chunk_ready(k, self.WORLD_SIZE, side()) if self._overlap_comm() else None)
                                      for k in (0, 1))
        return self._arena

    # Communication modes (WORLD_SIZE > 1).  Default = the round-1 scheme that the 1->8 GPU scaling runs were measured with: ONE
    # eager all-reduce of the arena between graph A and graph B, plus one eager broadcast of rank 0's BN statistics before the
    # step.  ETB_COMM_OVERLAP=1: the all-reduce is issued in 3 chunks from autograd marks during backward on a communication
    # stream (eager steps); ETB_COMM_IN_GRAPH=1 additionally captures the NCCL calls inside graph A (no host round-trip).
    COMM_OVERLAP = os.environ.get("ETB_COMM_OVERLAP", "0") == "1"
    COMM_IN_GRAPH = os.environ.get("ETB_COMM_IN_GRAPH", "0") == "1"

    def _overlap_comm(self):
        """the chunked all-reduce may run INSIDE backward only when every backward is followed by an optimizer step
        (accumulate == 1): gradients accumulate in the arena across iterations, so they must be reduced once per step"""
        if self.WORLD_SIZE <= 1 or not (self.fixed_accumulate or max(round(64 / self.batch_size), 1) == 1):
            return False
        if torch.cuda.is_current_stream_capturing():
            return self.COMM_IN_GRAPH
        return self.COMM_OVERLAP or self.COMM_IN_GRAPH

    # "sum" = the reference (loss * WORLD_SIZE, then DDP's mean: trainer/ssod_trainer.py:638-648).  "avg" (ncclAvg: the same
    # collective at the same cost) is for synthetic benchmarks only: with SUM the effective learning rate grows with the world
    # size, and a random-init model's BatchNorm scales then drift WORLD_SIZE x faster (bench.py's self-check explains why that
    # matters); set before the first step.
    GRAD_REDUCE = os.environ.get("ETB_GRAD_REDUCE", "sum")

    def _allreduce_grads(self):
        """WORLD_SIZE > 1: all-reduce of the gradient arena (the chunks that were not already issued during backward)"""
        if self.WORLD_SIZE <= 1:
            return
        self._arena.average = (self.GRAD_REDUCE == "avg")
        if self._arena._next == 0 and not (self.COMM_OVERLAP or self.COMM_IN_GRAPH):
            self._arena.all_reduce_sum(self.WORLD_SIZE)          # one collective over the whole arena
        else:
            self._arena.finish(self.WORLD_SIZE)

    def _bn_broadcast(self):
        """DDP broadcast_buffers=True: rank 0's BN running statistics overwrite every rank's before each forward"""
        if self.WORLD_SIZE > 1:
            if self._bn_sync is None:
                from .parallel import BnBufferSync
                self._bn_sync = BnBufferSync(self.model)
            self._bn_sync.broadcast(self.WORLD_SIZE)

    # trainer/ssod_trainer.py:458-488 (bf16 autocast needs no GradScaler; loss scale == 1), in three parts so that the
    # gradient all-reduce can sit between two captured CUDA graphs when WORLD_SIZE > 1
    WGRAD_SIDE_STREAM = os.environ.get("ETB_WGRAD_SIDE", "1") == "1"   # measured -0.9 ms/step (35.7 -> 34.9); ETB_WGRAD_SIDE=0 disables
    TEACHER_SIDE_STREAM = os.environ.get("ETB_TEACHER_SIDE", "1") == "1"   # teacher forward + NMS concurrent with the student forward

    def _backward(self, loss):
        self._ensure_arena()
        self._arena.begin_step()
        from . import autograd_conv as ac
        ac.backward(loss, side=self.WGRAD_SIDE_STREAM)   # weight-gradient branch on a side stream, joined before returning
        self._mark("backward")

    def _warmup(self, ni):
        """ssod_trainer.py:462-478: accumulate + per-iteration warm-up of lr / momentum (host scalars only; the fused SGD
        kernel reads them from device memory, so this also serves the captured step).  Group 2 -- the BatchNorm weights in
        the reference's group order (trainer.py:215-217) -- is the one that falls from warmup_bias_lr."""
        self.accumulate = 1 if self.fixed_accumulate else max(round(64 / self.batch_size), 1)
        if ni <= self.nw:
            xi = [0, self.nw]
            self.accumulate = max(1, np.interp(ni, xi, [1, 1 if self.fixed_accumulate else 64 / self.batch_size]).round())
            for j, x in enumerate(self.optimizer.param_groups):
                x['lr'] = float(np.interp(ni, xi, [self.warmup_bias_lr if j == 2 else 0.0, x['initial_lr'] * self.lf(self.epoch)]))
                if 'momentum' in x:
                    x['momentum'] = float(np.interp(ni, xi, [self.warmup_momentum, self.momentum]))
        return ni - self.last_opt_step >= self.accumulate

    def _step_and_ema(self, capturing=False):
        self.optimizer.step(zero_grad=True)      # fused SGD-Nesterov; also performs optimizer.zero_grad() on the arena
        if self.semi_ema:
            # == ema.update(model); semi_ema.update(ema.ema); inside a captured graph the decays come from device memory
            update_ema_pair(self.ema, self.semi_ema, self.model, scalars_dev=self._ema_scalars_dev if capturing else None)
        else:
            self.ema.update(self.model)

    def _optimizer_ema(self, ni):
        if self._warmup(ni):
            self._step_and_ema()
            self.last_opt_step = ni

    def update_optimizer(self, loss, ni):
        self._backward(loss)
        self._allreduce_grads()
        self._mark("allreduce")
        self._optimizer_ema(ni)

    def _mark(self, name):
        if self.profile:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.phase_events.append((name, ev))

    def phase_times_ms(self):
        """After a synchronize: {phase: total ms} accumulated over the recorded steps."""
        out = {}
        for (n0, e0), (n1, e1) in zip(self.phase_events[:-1], self.phase_events[1:]):
            if n1 != "start":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out

    def split_predict_and_feature(self, total_pred, total_feature, n_img):
        """ssod_trainer.py:568-585.  split_batch == t[:n], t[n:] whose backward is a no-op (the loss kernels write both
        gradients into one buffer) instead of autograd's zeros + copy + add per slice."""
        from .autograd_conv import split_batch
        fs = [split_batch(f, n_img) for f in total_feature]
        ps = [split_batch(p, n_img) for p in total_pred]
        return [a for a, _ in ps], [a for a, _ in fs], [b for _, b in ps], [b for _, b in fs]

    # trainer/ssod_trainer.py:587-680 (logging / meters excluded: rank-0 host bookkeeping)
    def train_instance(self, imgs, targets, unlabeled_imgs, unlabeled_imgs_ori, unlabeled_gt, unlabeled_M, ni,
                       host_pseudo_labels=False, _stop_after_backward=False):
        n_img = imgs.shape[0]
        self._mark("start")
        if self.WORLD_SIZE > 1 and (self.COMM_IN_GRAPH or not torch.cuda.is_current_stream_capturing()):
            self._bn_broadcast()         # (captured steps: train_instance_graphed issues it before the replay)
        # The teacher forward + NMS + pseudo-label transform feed nothing but the unsupervised loss, and the student forward
        # does not depend on them: with the device-resident pseudo labels they run on a side stream, concurrently with the
        # student forward (the teacher's batch-16 kernels leave SMs idle on the deep, small maps; the student's fill them),
        # and are joined right before ComputeStudentMatchLoss.  Inside a captured graph the fork/join become parallel branches.
        overlap = self.TEACHER_SIDE_STREAM and not host_pseudo_labels and not self.profile
        main = torch.cuda.current_stream(self.device)
        if overlap:
            if self._teacher_stream is None:
                self._teacher_stream = torch.cuda.Stream(self.device)
            fork = torch.cuda.Event()
            fork.record(main)
            self._teacher_stream.wait_event(fork)
        with torch.cuda.stream(self._teacher_stream if overlap else main):
            with torch.no_grad():
                (teacher_pred, train_out), teacher_feature = self.ema.ema(unlabeled_imgs_ori, augment=False)
            self._mark("teacher_forward")
            if hasattr(self.pseudo_label_creator, "update_device"):      # LabelMatch: ssod_trainer.py:616-617 (labeled-target histogram)
                self.pseudo_label_creator.update_device(targets)
                self.pseudo_label_creator.count += imgs.shape[0]
                self.pseudo_label_creator.pse_count += unlabeled_imgs.shape[0]
            if host_pseudo_labels:   # the reference's return contract: CPU float64 rows + flag (one D2H sync)
                unlabeled_targets, invalid_target_shape = self.pseudo_label_creator.create_pseudo_label_online_with_gt(
                    teacher_pred, unlabeled_imgs, unlabeled_M, unlabeled_imgs_ori, unlabeled_gt, self.RANK)
                n_dev = None
                if not invalid_target_shape:
                    unlabeled_targets = unlabeled_targets.to(self.device)
            else:                    # device-resident twin: no host sync between teacher and student
                h, w = unlabeled_imgs.shape[2:]
                unlabeled_targets, n_dev = self.pseudo_label_creator.create_pseudo_label_device(teacher_pred, unlabeled_M, h, w)
                invalid_target_shape = False
            self._mark("nms_pseudo_label")
            if overlap:
                join = torch.cuda.Event()
                join.record(self._teacher_stream)
                self._teacher_keep = (teacher_pred, train_out, teacher_feature)    # alive until the join below
        with torch.autocast("cuda", dtype=self.amp_dtype):
            # == self.model(torch.cat([imgs, unlabeled_imgs], 0)) (ssod_trainer.py:620-622): the native stem reads both
            # batches in place (uint8 from the loaders or fp32), so the concatenated fp32 image never exists
            total_pred, total_feature = self.model([imgs, unlabeled_imgs])
        self._mark("student_forward")
        sup_pred, sup_feature, un_sup_pred, un_sup_feature = self.split_predict_and_feature(total_pred, total_feature, n_img)
        sup_loss, sup_loss_items = self.compute_loss(sup_pred, targets)
        d_loss = self.domain_loss(sup_feature)
        t_loss = self.target_loss(un_sup_feature)
        if self.cfg.SSOD.with_da_loss:
            sup_loss = sup_loss + d_loss * self.da_loss_weights + t_loss * self.da_loss_weights
        else:
            sup_loss = sup_loss + d_loss * 0 + t_loss * 0
        if overlap:
            main.wait_event(join)        # pseudo labels ready
            self._teacher_keep = None
        if invalid_target_shape:
            un_sup_loss = torch.zeros(1, device=self.device)
            un_sup_loss_items = dict(ss_box=0, ss_obj=0, ss_cls=0)
        else:
            un_sup_loss, un_sup_loss_items = self.compute_un_sup_loss(un_sup_pred, unlabeled_targets, n_dev)
        # DDP: loss*WORLD_SIZE then gradient mean == plain SUM all-reduce of per-rank gradients (no scaling here)
        loss = sup_loss + un_sup_loss * self.cfg.SSOD.teacher_loss_weight
        self._mark("losses")
        if _stop_after_backward:         # captured graph A ends here; with accumulate == 1 it also contains the (overlapped) all-reduce
            self._backward(loss)
            if self._overlap_comm():
                self._allreduce_grads()
            return loss.detach()
        self.update_optimizer(loss, ni)
        self._mark("optimizer_ema")
        # logging values only -- detached, so that no reference to this step's autograd graph (and to the AccumulateGrad
        # nodes of the parameters, which are tied to the stream they were created on) survives the step
        det = lambda d: {k: (v.detach() if torch.is_tensor(v) else v) for k, v in d.items()}  # noqa: E731
        self.last = dict(loss=loss.detach(), sup=det(sup_loss_items), unsup=det(un_sup_loss_items))
        return loss.detach()

    # ---- the whole step as CUDA graphs ------------------------------------------------------------------------------
    def train_instance_graphed(self, imgs, targets, unlabeled_imgs, unlabeled_imgs_ori, unlabeled_gt, unlabeled_M, ni):
        """train_instance captured once (static shapes, device-resident pseudo labels, no host sync anywhere in the step)
        and replayed as two graphs: A = teacher forward ... backward (gradients accumulate in the arena), B = SGD-Nesterov +
        both EMA updates.  B is replayed on the iterations the reference's cadence steps the optimizer
        (ssod_trainer.py:462-488: `accumulate`, warm-up); for WORLD_SIZE > 1 the NCCL all-reduce sits between A and B.  The
        ~1.5k kernel launches + the autograd traversal collapse into two cudaGraphLaunch calls.  Inputs are copied into
        static buffers; learning rate / momentum (warm-up, scheduler) and the EMA decays of the step are host scalars
        written to device memory before B is replayed, so the schedule needs no re-capture."""
        if self.semi_ema is None:
            raise NotImplementedError("graphed step needs the fused ema/semi_ema pair (burn_epochs == 0)")
        shapes = (tuple(imgs.shape), tuple(targets.shape), tuple(unlabeled_imgs.shape), tuple(unlabeled_M.shape),
                  imgs.dtype, unlabeled_imgs.dtype, unlabeled_imgs_ori.dtype)      # uint8 loader batches vs fp32: different static buffers
        if self._graph is not None and self._graph["shapes"] != shapes:
            self.reset_graph()
        if self._graph is None:
            self._capture(imgs, targets, unlabeled_imgs, unlabeled_imgs_ori, unlabeled_M, ni, shapes)
        g = self._graph
        g["imgs"].copy_(imgs, non_blocking=True)
        g["targets"].copy_(targets, non_blocking=True)
        g["us"].copy_(unlabeled_imgs, non_blocking=True)
        g["uw"].copy_(unlabeled_imgs_ori, non_blocking=True)
        g["Ms"].copy_(unlabeled_M, non_blocking=True)
        comm_in_a = self.WORLD_SIZE > 1 and self.COMM_IN_GRAPH and (self.fixed_accumulate or max(round(64 / self.batch_size), 1) == 1)
        # everything the host contributes to this iteration is enqueued BEFORE graph A, so that A, the all-reduce and B follow
        # each other on the stream without a host gap: accumulate / lr / momentum of iteration ni (host scalars), and -- when
        # the optimizer is due -- the EMA decays and the SGD hyper-parameters (stream-ordered copies: the previous replay of B
        # has consumed the old values by the time they land)
        due = self._warmup(ni)
        if due:
            d1, d2 = next_pair_decays(self.ema, self.semi_ema)
            # pageable source: the runtime stages the 16 bytes before returning, so the next step cannot overwrite them early
            self._ema_scalars_dev.copy_(torch.tensor(ema_scalars(d1, d2), dtype=torch.float32))
            self.optimizer.refresh_hyper()   # lr / momentum of this step -> device memory read by the captured SGD kernel
        if self.WORLD_SIZE > 1 and not self.COMM_IN_GRAPH:
            self._bn_broadcast()
        g["graph"].replay()
        if due:
            if self.WORLD_SIZE > 1 and not comm_in_a:
                self._arena.begin_step()
                self._allreduce_grads()      # one SUM all-reduce per optimizer step, between the two graphs (enqueued, no host sync)
            g["graph_b"].replay()
            self.last_opt_step = ni
        if hasattr(self.pseudo_label_creator, "stage_detections"):   # LabelMatch: the captured step cannot stage its detections itself
            self.pseudo_label_creator.stage_detections()
        return g["loss"]

    def after_epoch(self, epoch, start_epoch=0):
        """ssod_trainer.py:319-323: LabelMatch re-estimates the per-class thresholds once per epoch; the unsupervised loss
        picks them up (and a captured step has to be re-captured because the thresholds are device constants of the graph)."""
        c = self.pseudo_label_creator
        if hasattr(c, "update_epoch_cls_thr") and epoch >= getattr(self.cfg.SSOD, "dynamic_thres_epoch", 0):
            c.update_epoch_cls_thr(epoch - start_epoch)
            self.compute_un_sup_loss.ignore_thres_high = list(c.cls_thr_high)
            self.compute_un_sup_loss.ignore_thres_low = list(c.cls_thr_low)
            self.reset_graph()

    def reset_graph(self):
        self._graph = None
        self._ema_scalars_dev = None

    def _capture(self, imgs, targets, us, uw, Ms, ni, shapes):
        dev = self.device
        st = dict(shapes=shapes, imgs=imgs.clone(), targets=targets.clone(), us=us.clone(), uw=uw.clone(),
                  Ms=Ms.to(dev, torch.float64).clone())
        self._ema_scalars_dev = torch.zeros(4, dtype=torch.float32, device=dev)
        was_profile, self.profile = self.profile, False
        self.last = {}
        # the warm-up steps below really train: snapshot every piece of state they touch and restore it afterwards
        self._ensure_arena()
        had_momentum = any(len(self.optimizer.state[p]) for g_ in self.optimizer.param_groups for p in g_["params"])
        tensors = [t for m in (self.model, self.ema.ema, self.semi_ema.ema) for t in m.state_dict().values()]
        if had_momentum:
            tensors += [self.optimizer.state[p]["momentum_buffer"] for g_ in self.optimizer.param_groups for p in g_["params"]
                        if self.optimizer.state[p].get("momentum_buffer") is not None]
        tensors.append(self._arena.flat)     # gradients already accumulated towards the next optimizer step (accumulate > 1)
        snap = [t.clone() for t in tensors]
        saved = (self.last_opt_step, self.ema.updates, self.semi_ema.updates, self.accumulate,
                 [(x['lr'], x.get('momentum')) for x in self.optimizer.param_groups])
        lm_state = None
        if hasattr(self.pseudo_label_creator, "count"):
            lm_state = (self.pseudo_label_creator.count, self.pseudo_label_creator.pse_count)
        # warm-up on a side stream (allocator + lazily-created state: momentum buffers, chunk tables, workspaces, TMA/func attrs)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self.train_instance(st["imgs"], st["targets"], st["us"], st["uw"], None, st["Ms"], ni, _stop_after_backward=True)
                if self._arena._next == 0:      # not already reduced inside train_instance (overlapped mode)
                    self._allreduce_grads()
                self._warmup(ni)
                self._step_and_ema()          # the optimizer + EMA branch is exercised (and later captured) unconditionally
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            st["loss"] = self.train_instance(st["imgs"], st["targets"], st["us"], st["uw"], None, st["Ms"], ni, _stop_after_backward=True)
        st["graph"] = graph
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb, pool=graph.pool()):
            self._step_and_ema(capturing=True)
        st["graph_b"] = gb
        with torch.no_grad():
            for t, c in zip(tensors, snap):
                t.copy_(c)
            if not had_momentum:      # buffers created by the warm-up: zero == "not yet created" for SGD (buf = grad on first use)
                for g_ in self.optimizer.param_groups:
                    for p in g_["params"]:
                        b = self.optimizer.state[p].get("momentum_buffer")
                        if b is not None:
                            b.zero_()
        self.last_opt_step, self.ema.updates, self.semi_ema.updates, self.accumulate, hyp = saved
        for x, (lr, mom) in zip(self.optimizer.param_groups, hyp):
            x['lr'] = lr
            if mom is not None:
                x['momentum'] = mom
        if lm_state is not None:
            self.pseudo_label_creator.count, self.pseudo_label_creator.pse_count = lm_state
        self.profile = was_profile
        self._graph = st


class SupTrainerStep:
    """The supervised step (trainer/trainer.py:406-443 train_in_epoch body + :381-404 update_optimizer), BASELINE configs
    #1/#2: student forward -> ComputeLoss -> backward -> [all-reduce] -> SGD-Nesterov -> ModelEMA.update, same native
    kernels as the SSOD step minus the teacher / pseudo-label path.  The optimizer cadence is the reference's:
    accumulate = max(round(64 / batch_size), 1) iterations per optimizer step (gradients add up in the arena in between)
    and the per-iteration warm-up of lr / momentum / accumulate while ni <= nw (trainer.py:372-376, 385-395)."""

    def __init__(self, cfg, device, rank=-1, world_size=1, epochs=None, batch_size=None, amp_dtype=torch.bfloat16, nb=None,
                 start_epoch=0):
        self.cfg, self.device = cfg, device
        self.RANK, self.WORLD_SIZE = rank, world_size
        self.epochs = epochs if epochs is not None else cfg.epochs
        self.epoch = 0
        self.batch_size = batch_size if batch_size is not None else cfg.Dataset.batch_size
        self.amp_dtype = amp_dtype
        self.model = SupModel(cfg).to(device)
        self.ema = ModelEMA(self.model)          # the reference keeps it on rank 0/-1 only (trainer.py:157); harmless elsewhere
        self.semi_ema = None
        self.fixed_accumulate = False            # trainer/trainer.py has no such switch
        SSODTrainerStep.build_optimizer(self, cfg)
        self.compute_loss = ComputeLoss(self.model, cfg)
        self._arena = None
        self.last_opt_step = -1
        if cfg.hyp.warmup_epochs > 0:
            self.nw = max(round(cfg.hyp.warmup_epochs * (nb or 0)), 1000)
            if nb:
                self.nw = min(self.nw, (self.epochs - start_epoch) / 2 * nb)
        else:
            self.nw = -1
        self._graph = None

    _warmup = SSODTrainerStep._warmup
    _ensure_arena = SSODTrainerStep._ensure_arena
    _bn_broadcast = SSODTrainerStep._bn_broadcast
    _bn_sync = None

    def _overlap_comm(self):          # the supervised step keeps the single all-reduce after backward
        return False

    def _forward_backward(self, imgs, targets):
        if self.WORLD_SIZE > 1 and not torch.cuda.is_current_stream_capturing():
            self._bn_broadcast()      # DDP broadcast_buffers=True (captured steps: issued before the replay)
        with torch.autocast("cuda", dtype=self.amp_dtype):
            pred = self.model(imgs)
        loss, loss_items = self.compute_loss(pred, targets)
        self._ensure_arena()
        from . import autograd_conv as ac
        ac.backward(loss, side=SSODTrainerStep.WGRAD_SIDE_STREAM)
        return loss.detach()

    def _step_and_ema(self):
        self.optimizer.step(zero_grad=True)
        self.ema.update(self.model)

    def train_step(self, imgs, targets, ni):
        loss = self._forward_backward(imgs, targets)
        if self._warmup(ni):
            self._arena.average = (SSODTrainerStep.GRAD_REDUCE == "avg")
            self._arena.all_reduce_sum(self.WORLD_SIZE)
            self._step_and_ema()
            self.last_opt_step = ni
        return loss

    def train_step_graphed(self, imgs, targets, ni):
        """train_step with forward + loss + backward replayed from one captured CUDA graph (static shapes); the optimizer /
        EMA launches (2 kernels, host-side decay) stay eager on the iterations the cadence asks for."""
        shapes = (tuple(imgs.shape), tuple(targets.shape), imgs.dtype)
        if self._graph is None or self._graph["shapes"] != shapes:
            self._ensure_arena()
            st = dict(shapes=shapes, imgs=imgs.clone(), targets=targets.clone())
            tensors = [t for t in self.model.state_dict().values()] + [self._arena.flat]
            snap = [t.clone() for t in tensors]
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._forward_backward(st["imgs"], st["targets"])
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["loss"] = self._forward_backward(st["imgs"], st["targets"])
            st["graph"] = graph
            with torch.no_grad():
                for t, c in zip(tensors, snap):       # BN running statistics / num_batches_tracked / the arena
                    t.copy_(c)
            self._graph = st
        g = self._graph
        g["imgs"].copy_(imgs, non_blocking=True)
        g["targets"].copy_(targets, non_blocking=True)
        self._bn_broadcast()
        g["graph"].replay()
        if self._warmup(ni):
            self._arena.average = (SSODTrainerStep.GRAD_REDUCE == "avg")
            self._arena.all_reduce_sum(self.WORLD_SIZE)
            self._step_and_ema()
            self.last_opt_step = ni
        return g["loss"]


class DevicePrefetcher:
    """Double-buffered host -> device staging of a batch on a side stream, so the H2D copy of step i+1 overlaps the kernels
    of step i (the reference's loop copies on the compute stream: `imgs.to(device, non_blocking=True)`,
    trainer/ssod_trainer.py:694-696).  put(batch of pinned host tensors) enqueues the copies into the next slot;
    get() makes the current stream wait for the oldest slot and returns its device tensors.  A slot is reused two put()s
    later, i.e. after the step that consumed it has been enqueued on the compute stream -- put() makes the copy stream wait
    for that point before overwriting."""

    def __init__(self, device, slots=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [dict(buf=None, ready=torch.cuda.Event(), free=None) for _ in range(slots)]
        self.head = self.tail = 0          # next slot to fill / next slot to hand out
        self.pending = 0

    def put(self, batch):
        assert self.pending < len(self.slots), "prefetcher full: call get() first"
        sl = self.slots[self.head]
        if sl["buf"] is None:
            sl["buf"] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in batch.items()}
        if sl["free"] is not None:
            self.stream.wait_event(sl["free"])          # the consumer of this slot's previous contents has been enqueued and finished
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                sl["buf"][k].copy_(v, non_blocking=True)
            sl["ready"].record(self.stream)
        self.head = (self.head + 1) % len(self.slots)
        self.pending += 1

    def get(self):
        assert self.pending > 0, "prefetcher empty: call put() first"
        sl = self.slots[self.tail]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(sl["ready"])
        self._last = sl
        self.tail = (self.tail + 1) % len(self.slots)
        self.pending -= 1
        return sl["buf"]

    def release(self):
        """call after the kernels that read the last get()'s tensors have been enqueued on the current stream"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._last["free"] = ev
