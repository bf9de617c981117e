"""Iterative Error Feedback regressor, B200-native (drop-in for reference models/ief_module.py:8-64).

Same constructor, attributes (fc1/fc2/fc3, relu, ief_layers -- which registers the three Linear layers a second
time, hence the duplicated state_dict keys -- iterations, initial_params_estimate) and return convention (three
column views of one [B,157] tensor).  forward() runs every iteration in a single cluster kernel (csrc/ief.cu).
"""
import numpy as np
import torch
import torch.nn as nn

import config
from straps_b200.engine import RegressorEngine, require_inference


class IEFModule(nn.Module):
    def __init__(self, fc_layers_neurons, in_features, num_output_params, iterations=3):
        super(IEFModule, self).__init__()
        if list(fc_layers_neurons) != [512, 512] or in_features != 512 or num_output_params != 157:
            raise NotImplementedError('the B200 IEF kernel is built for the ResNet-18 configuration (512 -> 157)')
        self.fc1 = nn.Linear(in_features + num_output_params, fc_layers_neurons[0])
        self.fc2 = nn.Linear(fc_layers_neurons[0], fc_layers_neurons[1])
        self.fc3 = nn.Linear(fc_layers_neurons[1], num_output_params)
        self.relu = nn.ReLU(inplace=True)
        for fc in (self.fc1, self.fc2, self.fc3):
            nn.init.zeros_(fc.bias)
        self.ief_layers = nn.Sequential(self.fc1, self.relu, self.fc2, self.relu, self.fc3)
        self.iterations = iterations
        self.initial_params_estimate = self.load_mean_params_6d_pose(config.SMPL_MEAN_PARAMS_PATH)
        self._engine = RegressorEngine(ief=self)

    def load_mean_params_6d_pose(self, mean_params_path):
        """[0.9, 0, 0 | mean 6-D pose (144) | mean shape (10)] as float32 -- reference lines 33-46."""
        mean_smpl = np.load(mean_params_path)
        est = np.zeros(3 + 24 * 6 + 10)
        est[0] = 0.9
        est[3:] = np.concatenate((mean_smpl['pose'], mean_smpl['shape']))
        return torch.from_numpy(est.astype(np.float32)).float()

    def forward(self, img_features):
        require_inference(self, 'IEFModule.forward')
        params = self._engine.ief_forward(img_features, self.iterations)
        return params[:, :3], params[:, 3:3 + 24 * 6], params[:, 3 + 24 * 6:]
