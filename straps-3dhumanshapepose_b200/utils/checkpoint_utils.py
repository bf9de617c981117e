"""Checkpoint helpers (drop-in for reference utils/checkpoint_utils.py:4-26 + the save dict of train/...:365-377).

SURVEY.md 8f row N3: the on-disk format either side of the hot path.  A reference `.tar` loads as is: the regressor and the
criterion keep the reference's `state_dict` keys, and straps_b200.parallel.DataParallelAdam reads/writes torch.optim.Adam's
`state_dict` layout (per-parameter exp_avg / exp_avg_sq) from/to its flat moment buckets.
"""
import numpy as np
import torch

CHECKPOINT_KEYS = ('epoch', 'best_epoch', 'best_epoch_val_metrics', 'model_state_dict', 'best_model_state_dict',
                   'optimiser_state_dict', 'criterion_state_dict')


def load_training_info_from_checkpoint(checkpoint, save_val_metrics):
    """-> (current_epoch, best_epoch, best_model_wts, best_epoch_val_metrics); metrics not tracked any more are dropped and
    newly tracked ones start at +inf, as in the reference."""
    best_metrics = dict(checkpoint['best_epoch_val_metrics'])
    best_metrics = {m: best_metrics.get(m, np.inf) for m in save_val_metrics}
    current_epoch = checkpoint['epoch'] + 1
    print('\nTraining information loaded from checkpoint.')
    print('Current epoch:', current_epoch)
    print('Best epoch val metrics from last training run:', best_metrics, ' - achieved in epoch:', checkpoint['best_epoch'])
    return current_epoch, checkpoint['best_epoch'], checkpoint['best_model_state_dict'], best_metrics


def save_checkpoint(path, epoch, best_epoch, best_epoch_val_metrics, regressor, best_model_wts, optimiser, criterion):
    """Writes the reference's checkpoint dict (train/train_synthetic_otf_rendering.py:368-377)."""
    torch.save({'epoch': epoch, 'best_epoch': best_epoch, 'best_epoch_val_metrics': best_epoch_val_metrics,
                'model_state_dict': regressor.state_dict(), 'best_model_state_dict': best_model_wts,
                'optimiser_state_dict': optimiser.state_dict(), 'criterion_state_dict': criterion.state_dict()}, path)


def resume_from_checkpoint(path, regressor, optimiser, criterion, map_location=None):
    """run_train.py:204-209."""
    checkpoint = torch.load(path, map_location=map_location, weights_only=False)
    regressor.load_state_dict(checkpoint['model_state_dict'])
    optimiser.load_state_dict(checkpoint['optimiser_state_dict'])
    criterion.load_state_dict(checkpoint['criterion_state_dict'])
    return checkpoint
