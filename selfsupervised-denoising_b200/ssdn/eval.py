"""DenoiserEvaluator: evaluate a trained model (weights ``.wt`` or ``.training`` snapshot) on padded full-size images
(reference: ssdn/ssdn/eval.py:19-127, train.py:243-259, 553-584).

The evaluation path is the forward half of the hot path at image size: images are reflect-padded on the right / bottom
to a square multiple of 32 (``NoisyDataset(pad_multiple=NoiseNetwork.input_wh_mul(), square=blindspot)``,
noise_wrapper.py:183-232), denoised by ``Denoiser.run_pipeline`` on the CUDA engine, cropped back to the original
top-left region and scored by PSNR against the clean image.  ``set_test_data(path)`` + ``evaluate()`` reads a folder or HDF5
test set as the reference does; any iterable of ``NoisyDataset`` batches can be evaluated too.  Image dumps and
TensorBoard are outside the hot path."""
from __future__ import annotations

from typing import Callable, Dict, Iterable

import torch

from ssdn.denoiser import Denoiser
from ssdn.train import DEFAULT_RUN_DIR, DenoiserTrainer


class DenoiserEvaluator(DenoiserTrainer):
    def __init__(self, target_path, runs_dir: str = DEFAULT_RUN_DIR, run_dir: str = None, device: str = None):
        super().__init__({}, runs_dir=runs_dir, run_dir=run_dir)
        state_dict = torch.load(target_path, map_location="cpu", weights_only=False) if isinstance(target_path, str) else target_path
        if "denoiser" in state_dict:
            self.load_state_dict(state_dict, device=device)
        else:
            self.denoiser = Denoiser.from_state_dict(state_dict, device=device)
        self.cfg = self.denoiser.cfg
        self.init_state()

    def evaluate(self, batches: Iterable = None, output_callback: Callable[[int, Dict], None] = None) -> Dict:
        if self.denoiser is None:
            raise RuntimeError("Denoiser not initialised for evaluation")
        return super().evaluate(batches, output_callback)      # batches None: the test set chosen with set_test_data(path)
