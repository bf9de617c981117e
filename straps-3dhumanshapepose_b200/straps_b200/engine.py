"""Keeps a RegressorHandle (packed weights + workspace on one device) in step with nn.Module parameters."""
import torch

from . import ops
from ._lib import StrapsError

_BN_FIELDS = ('weight', 'bias', 'running_mean', 'running_var')


class RegressorEngine(object):
    """Binds an encoder parameter container and/or an IEF parameter container to the CUDA library.

    Weights stay owned by PyTorch in the reference state_dict layout (OIHW fp32, nn.Linear [out,in]); the
    library keeps a packed copy that is refreshed whenever a parameter's version counter or storage moves.
    """

    def __init__(self, encoder=None, ief=None, conv_mode=None):
        self.encoder, self.ief = encoder, ief
        self.conv_mode = conv_mode or ops.DEFAULT_CONV_MODE
        self._handle = None
        self._sig = None
        self._train_steps = 0      # running statistics are updated through raw pointers: force a repack afterwards

    def _initial_estimate(self, device):
        """Device copy of `ief.initial_params_estimate` (a plain attribute, normally a CPU tensor: models/ief_module.py:31), cached by
        the source tensor's identity / version so that a forward neither copies nor synchronises (the previous per-call
        `.cpu()` comparison drained the stream once per step and made the forward impossible to capture in a CUDA graph)."""
        src = self.ief.initial_params_estimate
        key = (id(src), src.data_ptr(), src._version, str(device))
        if getattr(self, '_init_key', None) != key:
            self._init_dev = src.detach().to(device=device, dtype=torch.float32).contiguous().clone()
            self._init_key = key
        return self._init_dev

    # -- tensors in the order the C ABI expects
    def _tensors(self, device):
        h = self._handle
        if self.encoder is not None:
            mods = dict(self.encoder.named_modules())
            conv_w, bn = [], []
            for name in h.conv_names:
                conv = mods[name]
                # "conv1" -> "bn1", "layerX.Y.conv2" -> "layerX.Y.bn2", "...downsample.0" -> "...downsample.1"
                bn_name = name[:-1] + '1' if name.endswith('downsample.0') else name.replace('conv', 'bn')
                b = mods[bn_name]
                conv_w.append(conv.weight)
                bn.append(tuple(getattr(b, f) for f in _BN_FIELDS))
        else:
            z = lambda *s: torch.zeros(*s, device=device)
            conv_w, bn = [], []
            shapes = _conv_shapes(h.c_in)
            for name in h.conv_names:
                co, ci, k = shapes[name]
                conv_w.append(z(co, ci, k, k))
                bn.append((torch.ones(co, device=device), z(co), z(co), torch.ones(co, device=device)))
        if self.ief is not None:
            fc_w = [self.ief.fc1.weight, self.ief.fc2.weight, self.ief.fc3.weight]
            fc_b = [self.ief.fc1.bias, self.ief.fc2.bias, self.ief.fc3.bias]
            init = self._initial_estimate(device)
        else:
            fc_w = [torch.zeros(512, 669, device=device), torch.zeros(512, 512, device=device), torch.zeros(157, 512, device=device)]
            fc_b = [torch.zeros(512, device=device), torch.zeros(512, device=device), torch.zeros(157, device=device)]
            init = torch.zeros(157, device=device)
        return conv_w, bn, fc_w, fc_b, init

    def _sync(self, device, batch, c_in):
        device = torch.device(device)
        if device.type != 'cuda':
            raise StrapsError('the B200 regressor only runs on CUDA devices (got %s); there is no CPU fallback' % device)
        h = self._handle
        if h is None or h.device != device or h.max_batch < batch or h.c_in != c_in:
            self._handle = ops.RegressorHandle(device, c_in, max(batch, h.max_batch if h is not None else 0))
            self._sig = None
        conv_w, bn, fc_w, fc_b, init = self._tensors(device)
        flat = list(conv_w) + [t for q in bn for t in q] + fc_w + fc_b
        for t in flat:
            if t.device != device:
                raise StrapsError('regressor parameters live on %s but the input is on %s' % (t.device, device))
        sig = tuple((t.data_ptr(), t._version) for t in flat) + (init.data_ptr(), self._train_steps)
        if sig != self._sig:
            self._handle.load([t.detach() for t in conv_w], [[t.detach() for t in q] for q in bn],
                              [t.detach() for t in fc_w], [t.detach() for t in fc_b], init)
            self._sig = sig
        return self._handle

    def _c_in(self, x=None):
        if self.encoder is not None:
            return self.encoder.conv1.weight.shape[1]
        return 1

    def encoder_forward(self, x):
        h = self._sync(x.device, x.shape[0], self._c_in())
        return h.encoder_forward(x, self.conv_mode)

    def ief_forward(self, feat, iters):
        h = self._sync(feat.device, feat.shape[0], self._c_in())
        return h.ief_forward(feat, iters)

    def forward(self, x, iters):
        h = self._sync(x.device, x.shape[0], self._c_in())
        return h.forward(x, self.conv_mode, iters)

    def forward_from_labels(self, seg_labels, joints2d, table, half_size, iters):
        h = self._sync(seg_labels.device, seg_labels.shape[0], self._c_in())
        return h.forward_from_labels(seg_labels, joints2d, table, half_size, iters)

    def read_activation(self, name, batch):
        return self._handle.read_activation(name, batch)

    # ---- training path: torch.autograd.Function over the library's forward/backward kernels ----
    def _train_tensors(self, device):
        conv_w, bn, fc_w, fc_b, init = self._tensors(device)
        return conv_w, bn, fc_w, fc_b

    def forward_train(self, x, iters, want='params'):
        """want: 'params' (encoder + IEF), 'feat' (encoder only).  Differentiable wrt all 66 regressor parameters."""
        h = self._sync(x.device, x.shape[0], self._c_in())
        conv_w, bn, fc_w, fc_b = self._train_tensors(x.device)
        flat = list(conv_w) + [q[0] for q in bn] + [q[1] for q in bn] + list(fc_w) + list(fc_b)
        out = _RegressorTrain.apply(self, h, x, iters, want, *flat)
        self._after_batch_stat_forward()
        return out

    def forward_batch_stats(self, x, iters, want='params'):
        """Train-mode forward WITHOUT autograd (`regressor.train()` under torch.no_grad()): BatchNorm normalises with the batch
        statistics and updates the running ones, as nn.BatchNorm2d does whenever `.training` is set (reference
        models/resnet.py:201-216 under nn.Module.train()); nothing is recorded for a backward pass."""
        h = self._sync(x.device, x.shape[0], self._c_in())
        feat = h.encoder_train_forward(x, update_running_stats=True, mode=self.conv_mode)
        self._after_batch_stat_forward()
        return feat if want == 'feat' else h.ief_forward(feat, iters)

    def _after_batch_stat_forward(self):
        if self.encoder is not None:
            # the 20 int64 `num_batches_tracked` counters (nn.BatchNorm2d under .train()) in ONE launch
            torch._foreach_add_([m.num_batches_tracked for m in self.encoder.modules() if isinstance(m, torch.nn.BatchNorm2d)], 1)
        self._train_steps += 1


# positions, in _RegressorTrain's flat parameter order (20 conv weights, 20 BN weights, 20 BN biases, 3 fc weights, 3 fc biases), of the
# parameters whose gradients are final after blocks 7..6 of the backward pass: layer4's five convolutions and BatchNorms, and the IEF
_EARLY_PARAMS = tuple(range(15, 20)) + tuple(range(35, 40)) + tuple(range(55, 60)) + tuple(range(60, 66))
_EARLY_HOOKS = {}      # id(layer4.0.conv1.weight) -> callable; registered by DataParallelAdam.enable_overlap


class _RegressorTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, handle, x, iters, want, *params):
        feat = handle.encoder_train_forward(x, update_running_stats=True, mode=engine.conv_mode)
        ctx.handle, ctx.iters, ctx.want = handle, iters, want
        # the activations and batch statistics the backward needs stay in the handle's workspace, which EVERY later encoder forward
        # (train or inference) overwrites: remember which forward this graph belongs to
        ctx.generation = handle.generation
        ctx.params = params          # the nn.Parameters themselves: their gradient slots (straps_b200.parallel.FlatBucket) are looked up in backward
        ctx.conv_shapes = [tuple(p.shape) for p in params[:20]]
        ctx.bn_channels = [p.shape[0] for p in params[20:40]]
        if want == 'feat':
            ctx.save_for_backward(feat)
            return feat
        out, saved = handle.ief_forward_train(feat, iters)
        ctx.save_for_backward(feat, saved)
        return out

    @staticmethod
    def backward(ctx, g):
        h = ctx.handle
        if ctx.generation != h.generation:
            raise StrapsError('regressor backward: the encoder workspace has been overwritten by a later forward (generation %d, this '
                              'graph was recorded at %d).  The B200 regressor keeps ONE set of activations per device: call '
                              'backward() before the next forward of the same regressor (gradient accumulation = backward per '
                              'micro-batch)' % (h.generation, ctx.generation))
        # A parameter re-homed into a flat gradient bucket carries its slot (a view of the bucket); while it has no .grad yet the
        # kernels write the gradient straight into the slot and autograd adopts it as .grad -- no per-parameter accumulate launch
        # (71 of them per step before), and the bucket is complete the moment the backward kernels finish.
        slot = [getattr(p, '_straps_grad_slot', None) if (torch.is_tensor(p) and p.grad is None) else None for p in ctx.params]
        if ctx.want == 'feat':
            d_feat = g.contiguous()
            dfw = [None] * 3
            dfb = [None] * 3
        else:
            feat, saved = ctx.saved_tensors
            d_feat, dfw, dfb = h.ief_backward(feat, saved, g, ctx.iters, out_w=slot[60:63], out_b=slot[63:66])
        # A data-parallel optimiser (straps_b200.parallel.DataParallelAdam, world > 1) asks to be told when the deep end of the bucket is
        # complete: layer4 (convolutions 15..19) + IEF = 76 % of the gradient bytes after the first two blocks of the pass.  Its
        # all-reduce then runs on a side stream under the rest of the pass.
        early = _EARLY_HOOKS.get(id(ctx.params[15])) if ctx.want != 'feat' else None
        if early is not None and all(slot[i] is not None for i in _EARLY_PARAMS):
            dws, dbn = h.encoder_backward(d_feat, ctx.conv_shapes, ctx.bn_channels, out_w=slot[:20], out_bn=list(zip(slot[20:40], slot[40:60])),
                                          blocks=(7, 6, 1, 0))
            for i in _EARLY_PARAMS:
                ctx.params[i].grad = slot[i]
            early()
            dws, dbn = h.encoder_backward(d_feat, ctx.conv_shapes, ctx.bn_channels, blocks=(5, 0, 0, 1), dws=dws, dbn=dbn)
        else:
            dws, dbn = h.encoder_backward(d_feat, ctx.conv_shapes, ctx.bn_channels, out_w=slot[:20],
                                          out_bn=list(zip(slot[20:40], slot[40:60])))
        grads = list(dws) + [p[0] for p in dbn] + [p[1] for p in dbn] + list(dfw) + list(dfb)
        # Slot gradients are handed to .grad HERE and reported to autograd as None: AccumulateGrad would not adopt a view that other
        # references keep alive (it clones it -- measured on B200: 0 of 66 adopted), and skipping it also keeps the parameters'
        # AccumulateGrad nodes (and their streams) out of a CUDA-graph capture of the step.
        for i, (p, s) in enumerate(zip(ctx.params, slot)):
            if s is not None and grads[i] is not None:
                p.grad = s
                grads[i] = None
        return (None, None, None, None, None) + tuple(grads)


def needs_grad(module):
    return module.training and torch.is_grad_enabled()


def _conv_shapes(c_in):
    shapes = {'conv1': (64, c_in, 7)}
    cin = 64
    for L, cout in enumerate((64, 128, 256, 512)):
        for blk in range(2):
            p = 'layer%d.%d' % (L + 1, blk)
            shapes[p + '.conv1'] = (cout, cin, 3)
            shapes[p + '.conv2'] = (cout, cout, 3)
            if blk == 0 and L > 0:
                shapes[p + '.downsample.0'] = (cout, cin, 1)
            cin = cout
    return shapes


def require_inference(module, what):
    """Stand-alone halves (ResNet / IEFModule used on their own) only run the inference kernels."""
    if module.training and torch.is_grad_enabled():
        raise StrapsError('%s: the training path is wired through SingleInputRegressor (and ResNet for the encoder alone); '
                          'call .eval() or wrap the call in torch.no_grad() here' % what)
