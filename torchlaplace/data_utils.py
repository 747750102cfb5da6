"""``torchlaplace.data_utils.basic_collate_fn`` -- the one helper of the reference's ``data_utils`` module that
the hot path's callers import (examples/simple_demo.py:44, 277-305; reference definition:
torchlaplace/data_utils.py:51-74 with its helpers 154-187, 327-391).

Host-side batching, no arithmetic: a list of trajectories ``[T, d_obs]`` becomes the batch dictionary the training
loops read.  Re-implemented from the reference's behaviour (checked against it by
``tests/golden/make_golden.py`` -> ``tests/golden/collate.npz``); the rest of the reference's ``data_utils``
(VAE / ODE-RNN baselines, dataset normalisation) is outside this repository's scope.
"""
import torch

DEVICE = "cuda" if torch.cuda.is_available() else "cpu"


def _split(data, time_steps, extrap, dataset, observe_step, predict_step):
  """Observed / to-predict halves of a batch: interpolation keeps the whole window on both sides, extrapolation
  observes the first half (first third for ``hopper``) and predicts the rest, each subsampled by its step."""
  if not extrap:
    return dict(observed_data=data.clone(), observed_tp=time_steps.clone(), data_to_predict=data.clone(),
                tp_to_predict=time_steps.clone()), "interp"
  n_obs = data.size(1) // (3 if dataset == "hopper" else 2)
  return dict(observed_data=data[:, :n_obs, :][:, ::observe_step, :].clone(),
              observed_tp=time_steps[:, :n_obs][:, ::observe_step].clone(),
              data_to_predict=data[:, n_obs:, :][:, ::predict_step, :].clone(),
              tp_to_predict=time_steps[:, n_obs:][:, ::predict_step].clone()), "extrap"


def basic_collate_fn(batch, time_steps, extrap=False, data="", device=DEVICE, data_type="train", observe_step=1,
                     predict_step=1):  # pylint: disable=unused-argument
  """Collate ``batch`` (list of ``[T, d_obs]`` tensors) sampled at the shared ``time_steps [T]``.

  Returns ``{"observed_data", "observed_tp", "data_to_predict", "tp_to_predict", "observed_mask",
  "mask_predicted_data", "labels", "mode"}``; ``observed_tp`` / ``tp_to_predict`` have shape ``[1, T']``,
  ``observed_mask`` is all ones, ``mask_predicted_data`` and ``labels`` are ``None`` (``data_type`` and ``device``
  do not change the result, as in the reference)."""
  stacked = torch.stack(batch)
  out, mode = _split(stacked, time_steps.unsqueeze(0), extrap, data, observe_step, predict_step)
  out["observed_mask"] = torch.ones_like(out["observed_data"])
  out["mask_predicted_data"] = None
  out["labels"] = None
  out["mode"] = mode
  return out


__all__ = ["basic_collate_fn", "DEVICE"]
