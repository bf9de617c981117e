"""SingleInputRegressor, B200-native (drop-in for reference models/regressor.py:7-47).

proxy representation [B, C, 256, 256] -> (cam [B,3], pose6d [B,144], shape [B,10]); `.image_encoder` and
`.ief_module` keep the reference's names and state_dict (132 keys).  One engine packs both halves so the
whole forward is a straight run of library kernels on the caller's CUDA stream.
"""
import torch.nn as nn

from models.resnet import resnet18
from models.ief_module import IEFModule
from straps_b200.engine import RegressorEngine, needs_grad


class SingleInputRegressor(nn.Module):
    def __init__(self, resnet_in_channels=1, resnet_layers=18, ief_iters=3, conv_mode=None):
        """conv_mode: 'f16x3_tc' (tcgen05, default) or 'fp32_simt' (CUDA-core fp32) -- B200 extension."""
        super(SingleInputRegressor, self).__init__()
        if resnet_layers != 18:
            raise NotImplementedError('resnet_layers=%r: only the ResNet-18 regressor is on the B200 hot path'
                                      % (resnet_layers,))
        num_output_params = 3 + 24 * 6 + 10
        self.image_encoder = resnet18(in_channels=resnet_in_channels, pretrained=False)
        self.ief_module = IEFModule([512, 512], 512, num_output_params, iterations=ief_iters)
        self._engine = RegressorEngine(encoder=self.image_encoder, ief=self.ief_module, conv_mode=conv_mode)
        # the halves share the combined engine, so `.image_encoder(x)` / `.ief_module(f)` reuse the same packed weights
        self.image_encoder._engine = self._engine
        self.ief_module._engine = self._engine

    def forward(self, input):
        if needs_grad(self):
            # training step (reference train/...:186): batch-statistics BatchNorm, autograd through the library's backward
            params = self._engine.forward_train(input, self.ief_module.iterations)
        elif self.image_encoder.training:
            # train mode under torch.no_grad(): batch statistics + running-stat update, no graph (nn.BatchNorm2d semantics)
            params = self._engine.forward_batch_stats(input, self.ief_module.iterations)
        else:
            params = self._engine.forward(input, self.ief_module.iterations)
        return params[:, :3], params[:, 3:147], params[:, 147:]

    def forward_from_labels(self, seg_labels, joints2D, std=4):
        """B200 extension (inference): part-segmentation labels [B,256,256] + 2-D joints [B,J,2] -> (cam, pose, shape).

        Equals `self(torch.cat([convert_multiclass_to_binary_labels_torch(seg).unsqueeze(1),
        convert_2Djoints_to_gaussian_heatmaps_torch(joints2D, 256, std)], dim=1))` bit for bit -- the input assembly of
        reference train/train_synthetic_otf_rendering.py:178-182 -- but the 285 MB fp32 proxy representation is never written: the
        stem's input pack generates it on the fly, so only the labels and the joints have to reach the device."""
        from utils.label_conversions import _gaussian_table, _cuda_f32
        if needs_grad(self) or self.image_encoder.training:
            raise RuntimeError('forward_from_labels is an inference path: call .eval() and wrap the call in torch.no_grad()')
        seg = _cuda_f32(seg_labels, 'seg_labels')
        j2d = _cuda_f32(joints2D, 'joints2D')
        params = self._engine.forward_from_labels(seg, j2d, _gaussian_table(std, seg.device), 2 * std, self.ief_module.iterations)
        return params[:, :3], params[:, 3:147], params[:, 147:]
