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
            init = self.ief.initial_params_estimate.to(device=device, dtype=torch.float32)
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
        sig = tuple((t.data_ptr(), t._version) for t in flat) + (init.cpu().numpy().tobytes(),)
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

    def read_activation(self, name, batch):
        return self._handle.read_activation(name, batch)


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
    """The training path (batch-statistics BN + backward kernels) is not part of this round."""
    if module.training and torch.is_grad_enabled():
        raise StrapsError('%s: train-mode forward/backward kernels are not built in this round; call .eval() '
                          'or wrap the call in torch.no_grad() (eval-mode BatchNorm statistics are used)' % what)
