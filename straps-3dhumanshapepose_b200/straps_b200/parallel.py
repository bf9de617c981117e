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
        with torch.no_grad():
            for p in plist:
                n = p.numel()
                self.params[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.params[o:o + n].view(p.shape)
                p.grad = self.grads[o:o + n].view(p.shape)
                self.offsets.append(o)
                o += n

    def zero_grad(self):
        self.grads.zero_()
        for p, o in zip(self.plist, self.offsets):     # a set_to_none zero_grad elsewhere must not detach the views
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * o:
                p.grad = self.grads[o:o + p.numel()].view(p.shape)

    def bump_versions(self):
        """The fused optimiser writes through the flat buffer: tell the engines that the weights changed."""
        with torch.no_grad():
            for p in self.plist:
                p.add_(0)


class DataParallelAdam(object):
    """all-reduce(flat grads) + fused Adam on the flat parameter bucket (defaults of run_train.py:200-201)."""

    def __init__(self, parameters, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, process_group=None, broadcast=True):
        self.bucket = FlatBucket(parameters)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.group = process_group
        self.exp_avg = torch.zeros_like(self.bucket.params)
        self.exp_avg_sq = torch.zeros_like(self.bucket.params)
        self.step_count = 0
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        if broadcast and self.world > 1:
            dist.broadcast(self.bucket.params, src=0, group=process_group)     # identical replicas

    def zero_grad(self):
        self.bucket.zero_grad()

    def all_reduce(self):
        """THE one collective of the step: sum the flat gradient bucket over ranks."""
        if self.world > 1:
            dist.all_reduce(self.bucket.grads, op=dist.ReduceOp.SUM, group=self.group)

    def step(self):
        self.all_reduce()
        self.step_count += 1
        b = self.bucket
        if b.params.is_cuda:
            ops.adam_step(b.params, b.grads, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr, self.betas, self.eps,
                          grad_scale=1.0 / self.world)
        else:
            # host-side restatement used by the gloo CPU tests of the bucket / collective logic only
            g = b.grads / self.world
            self.exp_avg.mul_(self.betas[0]).add_(g, alpha=1 - self.betas[0])
            self.exp_avg_sq.mul_(self.betas[1]).addcmul_(g, g, value=1 - self.betas[1])
            bc1 = 1 - self.betas[0] ** self.step_count
            bc2 = 1 - self.betas[1] ** self.step_count
            denom = self.exp_avg_sq.sqrt() / (bc2 ** 0.5) + self.eps
            b.params.addcdiv_(self.exp_avg, denom, value=-self.lr / bc1)
        b.bump_versions()


    # ---- checkpoint format of torch.optim.Adam (reference train/...:365-377, run_train.py:204-209) ----
    def state_dict(self):
        """The dict torch.optim.Adam.state_dict() would give for the same parameter list (per-parameter exp_avg / exp_avg_sq
        views of the flat moments), so `checkpoint['optimiser_state_dict']` interchanges with the reference's files."""
        b = self.bucket
        state = {}
        if self.step_count > 0:
            for i, p, o in zip(b.index, b.plist, b.offsets):       # frozen parameters have no entry, as in torch.optim.Adam
                n = p.numel()
                state[i] = {'step': torch.tensor(float(self.step_count)),
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
