"""One K1 launch on 8192 'L' images of 512x512 (an `ncu` target)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "kobato-eyes_b200"))
import torch
from kobato_b200 import ops
bank = ops.synth_images_device(0, 8192, 512, 512, 1, n_set=8192)
ops.phash_dhash_batch(bank[:64]); ops.phash_dhash_batch(bank); torch.cuda.synchronize()
