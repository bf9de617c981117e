"""Training / validation loss and metric tracker with device-resident sums.

Drop-in for the reference's metrics/train_loss_and_metrics_tracker.py:8-275: same constructor, same history keys, same
pickle log, same per-epoch normalisation.  What changed is `update_per_batch`: the reference copies four [B,6890,3]
tensors and every prediction to the host, aligns sample by sample in numpy and calls `.item()` on the losses -- the
largest synchronisation left in a training step.  Here every sum is accumulated by library kernels (csrc/metrics.cu)
into one float64 CUDA vector and read back ONCE per epoch (`sync()` / `update_per_epoch()`); a step never waits for
its metrics.  `loss_metric_sums` keeps the reference's dict-of-floats form and is refreshed by `sync()`.
"""
import pickle

import torch

from straps_b200 import ops
from straps_b200._lib import StrapsError

_SPLITS = ('train', 'val')
_TASKS = ('verts', 'shape_params', 'pose_params', 'joints2D', 'joints3D')
# three consecutive slots per point set: plain, scale+translation corrected, Procrustes aligned (the kernel's output order)
_POINT_SETS = (('pves', 'pves_sc', 'pves_pa'), ('pve-ts', 'pve-ts_sc', 'pve-ts_pa'), ('mpjpes', 'mpjpes_sc', 'mpjpes_pa'))
_ROW_METRICS = ('pose_mses', 'shape_mses', 'joints2D_l2es')
_METRICS = tuple(m for group in _POINT_SETS for m in group) + _ROW_METRICS
# values per sample that turn a sum into a mean (reference update_per_epoch)
_PER_SAMPLE = (('pve', 6890), ('mpjpe', 14), ('joints2D', 17), ('shape_mse', 10), ('pose_mse', 24 * 3 * 3))


def _sum_keys(split):
    return [split + '_losses'] + ['%s_%s_losses' % (split, t) for t in _TASKS] + ['%s_%s' % (split, m) for m in _METRICS]


class TrainingLossesAndMetricsTracker:
    def __init__(self, losses_to_track, metrics_to_track, img_wh, log_path, load_logs=False, current_epoch=None):
        self.all_per_task_loss_types = ['%s_%s_losses' % (s, t) for t in _TASKS for s in _SPLITS]
        self.all_metrics_types = ['%s_%s' % (s, m) for m in _METRICS for s in _SPLITS]
        self.losses_to_track = losses_to_track
        self.metrics_to_track = metrics_to_track
        self.img_wh = img_wh
        self.log_path = log_path
        if load_logs:
            self.history = self.load_history(log_path, current_epoch)
        else:
            self.history = {k: [] for k in ['train_losses', 'val_losses'] + self.all_per_task_loss_types + self.all_metrics_types}
        self.loss_metric_sums = None
        self._slot = {k: i for i, k in enumerate(_sum_keys('train') + _sum_keys('val'))}
        self._dev_sums = None          # float64 CUDA vector, one slot per key of self._slot
        print('Metrics tracker initialised.')

    # ------------------------------------------------------------------ history on disk
    def load_history(self, load_log_path, current_epoch):
        """Resume: read the pickle log, cut every series to `current_epoch` entries, zero-fill series the log lacks."""
        with open(load_log_path, 'rb') as f:
            history = pickle.load(f)
        for key in ['train_losses', 'val_losses'] + self.all_per_task_loss_types + self.all_metrics_types:
            if key in history:
                history[key] = history[key][:current_epoch]
            else:
                history[key] = [0.0] * current_epoch
                print(key, 'filled with zeros up to epoch', current_epoch)
        for key, series in history.items():
            assert len(series) == current_epoch, \
                "{} elements in {} list when current epoch is {}".format(len(series), key, current_epoch)
        print('Logs loaded from', load_log_path)
        return history

    # ------------------------------------------------------------------ per-epoch sums
    def initialise_loss_metric_sums(self):
        self.loss_metric_sums = {k: 0. for k in self._slot}
        self.loss_metric_sums['train_num_samples'] = 0
        self.loss_metric_sums['val_num_samples'] = 0
        if self._dev_sums is not None:
            self._dev_sums.zero_()

    def _sums_on(self, device):
        if self._dev_sums is None:
            self._dev_sums = torch.zeros(len(self._slot), dtype=torch.float64, device=device)
        elif self._dev_sums.device != device:
            raise StrapsError('metrics tracker: batches arrive on %s but the sums live on %s' % (device, self._dev_sums.device))
        return self._dev_sums

    def sync(self):
        """One device->host copy of all sums into `loss_metric_sums` (the only synchronisation of the tracker)."""
        if self._dev_sums is not None:
            host = self._dev_sums.cpu().tolist()
            for key, i in self._slot.items():
                self.loss_metric_sums[key] = host[i]
        return self.loss_metric_sums

    def update_per_batch(self, split, loss, task_losses_dict, pred_dict, target_dict, num_inputs_in_batch,
                         pred_reposed_vertices=None, target_reposed_vertices=None):
        assert split in _SPLITS, "Invalid split in metric tracker batch update."
        track = self.metrics_to_track
        if any('pve-ts' in m for m in track):
            assert (pred_reposed_vertices is not None) and (target_reposed_vertices is not None), \
                "Need to pass reposed vertices to metric tracker batch update."
        device = next((v.device for v in list(pred_dict.values()) + [loss] if torch.is_tensor(v) and v.is_cuda), None)
        if device is None:
            raise StrapsError('metrics tracker: predictions must be CUDA tensors -- the B200 path has no CPU fallback')
        sums = self._sums_on(device)
        slot = lambda name: self._slot[split + '_' + name]
        f32 = lambda t: t.detach().float()

        # losses: (total, five tasks) * batch size, one launch
        scalars = [loss] + [task_losses_dict[t] if t in self.losses_to_track else 0. for t in _TASKS]
        scalars = [f32(v).reshape(()).to(device) if torch.is_tensor(v) else torch.tensor(float(v), device=device) for v in scalars]
        ops.accumulate(torch.stack(scalars), num_inputs_in_batch, sums[slot('losses'):])
        self.loss_metric_sums[split + '_num_samples'] += num_inputs_in_batch

        # point-set metrics: plain / scale-corrected / Procrustes sums of one (pred, target) pair in ONE launch
        pairs = ((_POINT_SETS[0], lambda: (pred_dict['verts'], target_dict['verts'])),
                 (_POINT_SETS[1], lambda: (pred_reposed_vertices, target_reposed_vertices)),
                 (_POINT_SETS[2], lambda: (pred_dict['joints3D'], target_dict['joints3D'])))
        for names, get in pairs:
            which = sum(bit for bit, n in zip((ops.METRIC_PLAIN, ops.METRIC_SC, ops.METRIC_PA), names) if n in track)
            if which:
                p, t = get()
                ops.points_metrics(f32(p), f32(t), which, sums=sums[slot(names[0]):])
        if 'pose_mses' in track:
            ops.rows_metric(f32(pred_dict['pose_params_rot_matrices']), f32(target_dict['pose_params_rot_matrices']),
                            sums[slot('pose_mses'):], 1, squared=True)
        if 'shape_mses' in track:
            ops.rows_metric(f32(pred_dict['shape_params']), f32(target_dict['shape_params']), sums[slot('shape_mses'):], 1, squared=True)
        if 'joints2D_l2es' in track:
            # prediction back from [-1,1] to pixels (utils/joints2d_utils.py:5-10), then the per-joint L2 distance
            ops.rows_metric(f32(pred_dict['joints2D']), f32(target_dict['joints2D']), sums[slot('joints2D_l2es'):], 2,
                            pred_add=1.0, pred_mul=self.img_wh / 2.0)

    def update_per_epoch(self):
        sums = self.sync()
        for split in _SPLITS:
            self.history[split + '_losses'].append(sums[split + '_losses'] / sums[split + '_num_samples'])
        for loss_type in self.all_per_task_loss_types:
            split, task = loss_type.split('_', 1)
            task = task[:-len('_losses')]
            if task in self.losses_to_track:
                self.history[loss_type].append(sums[loss_type] / sums[split + '_num_samples'])
            else:
                self.history[loss_type].append(0.)
        for metric_type in self.all_metrics_types:
            split, metric = metric_type.split('_', 1)
            if metric in self.metrics_to_track:
                per_sample = next(n for tag, n in _PER_SAMPLE if tag in metric_type)
                self.history[metric_type].append(sums[metric_type] / (sums[split + '_num_samples'] * per_sample))

        print('Finished epoch.')
        print('Train Loss: {:.5f}, Val Loss: {:.5f}'.format(self.history['train_losses'][-1], self.history['val_losses'][-1]))
        for metric in self.metrics_to_track:
            print('Train {}: {:.5f}, Val {}: {:.5f}'.format(metric, self.history['train_' + metric][-1],
                                                            metric, self.history['val_' + metric][-1]))
        with open(self.log_path, 'wb') as f_out:
            pickle.dump(self.history, f_out)

    def determine_save_model_weights_this_epoch(self, save_val_metrics, best_epoch_val_metrics):
        """True unless a tracked validation metric got worse than the best epoch so far."""
        return not any(self.history['val_' + m][-1] > best_epoch_val_metrics[m] for m in save_val_metrics)
