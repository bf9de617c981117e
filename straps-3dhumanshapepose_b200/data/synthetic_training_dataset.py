"""SMPL pose/shape parameter dataset (drop-in for reference data/synthetic_training_dataset.py:6-57)."""
import numpy as np
import torch
from torch.utils.data import Dataset

_SOURCES = ('h36m', 'up3d', '3dpw')


class SyntheticTrainingDataset(Dataset):
    """npz with `fnames`, `poses` [N,72], `shapes` [N,10] -> items {'pose': f32[72], 'shape': f32[10]}."""

    def __init__(self, npz_path, params_from='all'):
        assert params_from in ['all', 'h36m', 'up3d', '3dpw', 'not_amass']
        data = np.load(npz_path)
        self.fnames, self.poses, self.shapes = data['fnames'], data['poses'], data['shapes']
        if params_from != 'all':
            wanted = _SOURCES if params_from == 'not_amass' else (params_from,)
            keep = [i for i, name in enumerate(self.fnames) if str(name).startswith(wanted)]
            self.fnames = [self.fnames[i] for i in keep]
            self.poses = [self.poses[i] for i in keep]
            self.shapes = [self.shapes[i] for i in keep]

    def __len__(self):
        return len(self.poses)

    def __getitem__(self, index):
        if torch.is_tensor(index):
            index = index.tolist()
        pose, shape = self.poses[index], self.shapes[index]
        assert pose.shape == (72,) and shape.shape == (10,), \
            "Poses and shapes are wrong: {}, {}, {}".format(self.fnames[index], pose.shape, shape.shape)
        return {'pose': torch.from_numpy(pose.astype(np.float32)),
                'shape': torch.from_numpy(shape.astype(np.float32))}
