"""Batch-sharded data parallelism for the training step (BASELINE config 4).

The reference is single-process / single-GPU (run_train.py:23-26) and has no distributed code, so this is new:
one process per GPU (torchrun), every rank holds a full replica and runs the step on its own shard of the batch,
then ONE all-reduce (NCCL over NVLink / NVSwitch; gloo in the CPU tests) sums a single flat fp32 gradient bucket
-- 11,906,658 elements at C=17: encoder + IEF + the five loss log-variances -- and one fused Adam kernel updates the
flat parameter bucket with the 1/world_size factor folded in.  BatchNorm statistics stay per replica (the reference
has no SyncBN); replicas start from identical weights (broadcast from rank 0).
"""
import torch
import torch.distributed as dist

from . import ops


class FlatBucket(object):
    """Re-homes a list of parameters into ONE contiguous fp32 buffer (and their .grad into another).

    After construction `p.data` and `p.grad` of every parameter are views of `self.params` / `self.grads`, so autograd
    accumulates straight into the bucket and the optimiser / all-reduce work on two flat tensors.
    """

    def __init__(self, parameters):
        # `all_params` keeps EVERY parameter handed in, frozen ones included, because torch.optim.Adam numbers its state by position
        # in that list (run_train.py:200-201 passes regressor.parameters() + criterion.parameters(), i.e. also the log-variances of
        # switched-off losses, requires_grad=False: losses/multi_task_loss.py:47-71).  Only trainable ones live in the flat buffers.
        seen, self.all_params = set(), []
        for p in parameters:
            if id(p) not in seen:
                seen.add(id(p))
                self.all_params.append(p)
        self.index = [i for i, p in enumerate(self.all_params) if p.requires_grad]      # optimiser-state index of plist[k]
        plist = [self.all_params[i] for i in self.index]
        if not plist:
            raise ValueError('FlatBucket needs at least one parameter that requires grad')
        dev = plist[0].device
        self.plist = plist
        self.numel = sum(p.numel() for p in plist)
        self.params = torch.empty(self.numel, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.offsets = []
        o = 0
        self.slots = []
        with torch.no_grad():
            for p in plist:
                n = p.numel()
                self.params[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.params[o:o + n].view(p.shape)
                slot = self.grads[o:o + n].view(p.shape)
                p.grad = slot
                # the library's backward kernels write a fresh gradient straight into this view (straps_b200.engine._RegressorTrain)
                p._straps_grad_slot = slot
                self.slots.append(slot)
                self.offsets.append(o)
                o += n

    def zero_grad(self):
        """One memset of the bucket; every .grad is dropped (set-to-none), so the next backward WRITES its gradients into the slots
        instead of accumulating into them (gather() below puts anything that arrived some other way into place)."""
        self.grads.zero_()
        for p in self.plist:
            p.grad = None

    def gather(self):
        """After backward: make the flat bucket hold every gradient and every .grad a view of it.  Gradients the library wrote into
        their slot (and autograd adopted) are already there; anything else -- a tensor autograd allocated itself, an accumulated
        gradient, a set_to_none zero_grad by foreign code -- is copied in."""
        base = self.grads.data_ptr()
        for p, o, slot in zip(self.plist, self.offsets, self.slots):
            g = p.grad
            if g is None:
                p.grad = slot                   # no gradient this step: the slot is zero since zero_grad()
            elif g.data_ptr() != base + 4 * o:
                slot.copy_(g)
                p.grad = slot

    def bump_versions(self):
        """The fused optimiser writes through the flat buffer: tell the engines that the weights changed (version counters only --
        round 1 did this with 67 `p.add_(0)` launches that re-wrote all 47.6 MB of parameters)."""
        torch.autograd.graph.increment_version(self.plist)


class DataParallelAdam(object):
    """all-reduce(flat grads) + fused Adam on the flat parameter bucket (defaults of run_train.py:200-201)."""

    def __init__(self, parameters, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, process_group=None, broadcast=True):
        self.bucket = FlatBucket(parameters)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.group = process_group
        self.exp_avg = torch.zeros_like(self.bucket.params)
        self.exp_avg_sq = torch.zeros_like(self.bucket.params)
        # the step count lives on the device for CUDA buckets (bias corrections inside the kernel: the step is CUDA-graph capturable)
        self._step_host = 0
        self._step_dev = torch.zeros(1, dtype=torch.int64, device=self.bucket.params.device) if self.bucket.params.is_cuda else None
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if broadcast and self.world > 1:
            dist.broadcast(self.bucket.params, src=0, group=process_group)     # identical replicas

    @property
    def step_count(self):
        return int(self._step_dev.item()) if self._step_dev is not None else self._step_host

    @step_count.setter
    def step_count(self, v):
        if self._step_dev is not None:
            self._step_dev.fill_(int(v))
        self._step_host = int(v)

    def zero_grad(self, set_to_none=True):
        self.bucket.zero_grad()

    # ---- overlap of the collective with the backward pass -------------------------------------------------------------------------
    _early_range = None      # [lo, hi) of the bucket that is complete after the first two blocks of the encoder's backward pass
    _early_pending = False   # its all-reduce has been started (side stream) during the current backward pass
    _early_in_graph = False  # ... and was captured into the step's first CUDA graph (GraphedTrainStep): replays need only the rest
    _side = None
    _hook_key = None

    def enable_overlap(self, regressor):
        """Start the all-reduce of layer4 + IEF -- 9.1 M of the 11.9 M gradients, final after the first two blocks of the encoder's
        backward pass -- on a side stream while the remaining six blocks and the stem run (straps_b200.engine._RegressorTrain splits
        its pass there and calls back); `step()` then reduces only the rest.  Returns False (and changes nothing) when it does not
        apply: one rank, CPU buckets, or a parameter order in which that range is not contiguous."""
        from . import engine
        b = self.bucket
        if self.world <= 1 or not b.params.is_cuda:
            return False
        rng = self.early_range_of(regressor)
        if rng is None:
            return False
        self._early_range = rng
        self._side = torch.cuda.Stream(device=b.params.device)
        self._hook_key = id(regressor.image_encoder.layer4[0].conv1.weight)
        engine._EARLY_HOOKS[self._hook_key] = self._on_early_grads
        return True

    def early_range_of(self, regressor):
        """[lo, hi) of the flat bucket that holds exactly the parameters of layer4 and of the IEF module, or None if they are not one
        contiguous run of it (they are for `list(regressor.parameters()) + list(criterion.parameters())`, the reference's order)."""
        b = self.bucket
        try:
            early = list(regressor.image_encoder.layer4.parameters()) + list(regressor.ief_module.parameters())
        except AttributeError:
            return None
        ids = {id(p) for p in early if p.requires_grad}
        pos = [k for k, p in enumerate(b.plist) if id(p) in ids]
        if not pos or len(pos) != len(ids) or pos != list(range(pos[0], pos[0] + len(pos))):
            return None
        return (b.offsets[pos[0]], b.offsets[pos[-1]] + b.plist[pos[-1]].numel())

    def disable_overlap(self):
        from . import engine
        if self._hook_key is not None:
            engine._EARLY_HOOKS.pop(self._hook_key, None)
        self._early_range, self._hook_key, self._early_pending, self._early_in_graph = None, None, False, False

    def _on_early_grads(self):
        """Called from inside the backward pass (autograd's thread, on the stream of the forward) once bucket[lo:hi) is final."""
        lo, hi = self._early_range
        self._side.wait_stream(torch.cuda.current_stream(self.bucket.params.device))
        with torch.cuda.stream(self._side):
            dist.all_reduce(self.bucket.grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        self._early_pending = True

    def _join_early(self):
        torch.cuda.current_stream(self.bucket.params.device).wait_stream(self._side)
        self._early_pending = False

    def exchange(self):
        """THE collective of the step: sum the flat gradient bucket over ranks (NCCL all-reduce over NVLink / NVSwitch) -- all of it,
        or what the early all-reduce of this step has not covered."""
        if self.world <= 1:
            return
        g = self.bucket.grads
        if self._early_pending or self._early_in_graph:
            lo, hi = self._early_range
            if lo > 0:
                dist.all_reduce(g[:lo], op=dist.ReduceOp.SUM, group=self.group)
            if hi < g.numel():
                dist.all_reduce(g[hi:], op=dist.ReduceOp.SUM, group=self.group)
            if self._early_pending:
                self._join_early()
        else:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce(self):
        self.bucket.gather()
        self.exchange()

    _deferred = False        # set by straps_b200.graphs.GraphedTrainStep while it captures the step in two graphs around the collective

    def step(self):
        if self._deferred:
            self.bucket.gather()             # captured with the backward; exchange() + apply_update() are driven by the graph wrapper
            if self._early_pending:          # the side stream's all-reduce is part of this capture: join it before the capture ends
                self._join_early()
                self._early_in_graph = True
            return
        self.all_reduce()
        self.apply_update()

    def apply_update(self):
        b = self.bucket
        if b.params.is_cuda:
            ops.adam_step_dev(b.params, b.grads, self.exp_avg, self.exp_avg_sq, self._step_dev, self.lr, self.betas, self.eps,
                              grad_scale=1.0 / self.world)
        else:
            self._step_host += 1
            # host-side restatement used by the gloo CPU tests of the bucket / collective logic only
            g = b.grads / self.world
            self.exp_avg.mul_(self.betas[0]).add_(g, alpha=1 - self.betas[0])
            self.exp_avg_sq.mul_(self.betas[1]).addcmul_(g, g, value=1 - self.betas[1])
            bc1 = 1 - self.betas[0] ** self._step_host
            bc2 = 1 - self.betas[1] ** self._step_host
            denom = self.exp_avg_sq.sqrt() / (bc2 ** 0.5) + self.eps
            b.params.addcdiv_(self.exp_avg, denom, value=-self.lr / bc1)
        b.bump_versions()


    # ---- checkpoint format of torch.optim.Adam (reference train/...:365-377, run_train.py:204-209) ----
    def state_dict(self):
        """The dict torch.optim.Adam.state_dict() would give for the same parameter list (per-parameter exp_avg / exp_avg_sq
        views of the flat moments), so `checkpoint['optimiser_state_dict']` interchanges with the reference's files."""
        b = self.bucket
        state = {}
        step_count = self.step_count
        if step_count > 0:
            for i, p, o in zip(b.index, b.plist, b.offsets):       # frozen parameters have no entry, as in torch.optim.Adam
                n = p.numel()
                state[i] = {'step': torch.tensor(float(step_count)),
                            'exp_avg': self.exp_avg[o:o + n].view(p.shape).clone(),
                            'exp_avg_sq': self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {'lr': self.lr, 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': 0, 'amsgrad': False, 'maximize': False,
                 'foreach': None, 'capturable': False, 'differentiable': False, 'fused': None, 'decoupled_weight_decay': False,
                 'params': list(range(len(b.all_params)))}
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd):
        b = self.bucket
        groups = sd['param_groups']
        if len(groups) != 1 or len(groups[0]['params']) != len(b.all_params):
            raise ValueError('optimiser state has %d parameter groups / %d parameters, expected 1 / %d'
                             % (len(groups), len(groups[0]['params']) if groups else 0, len(b.all_params)))
        g = groups[0]
        if g.get('weight_decay', 0) or g.get('amsgrad', False):
            raise ValueError('weight decay / amsgrad are not part of the reference configuration (run_train.py:200-201)')
        self.lr, self.betas, self.eps = g['lr'], tuple(g['betas']), g['eps']
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        trainable = set(b.index)
        for k in sd['state']:
            if int(k) not in trainable:
                raise ValueError('optimiser state holds moments for parameter %s, which is frozen (requires_grad=False) here: the '
                                 'checkpoint was trained with another `losses_on`' % (k,))
        for i, p, o in zip(b.index, b.plist, b.offsets):
            st = sd['state'].get(i, sd['state'].get(str(i)))
            if st is None:
                continue
            n = p.numel()
            if tuple(st['exp_avg'].shape) != tuple(p.shape):
                raise ValueError('optimiser state %d has shape %s, parameter has %s' % (i, tuple(st['exp_avg'].shape), tuple(p.shape)))
            self.exp_avg[o:o + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise ValueError('per-parameter step counts differ (%s): not a plain Adam run' % sorted(steps))
        self.step_count = steps.pop() if steps else 0


def shard(batch_tensor, rank, world):
    """Contiguous batch shard of rank `rank`."""
    n = batch_tensor.shape[0]
    if n % world != 0:
        raise ValueError('global batch %d is not divisible by the world size %d' % (n, world))
    per = n // world
    return batch_tensor[rank * per:(rank + 1) * per]
