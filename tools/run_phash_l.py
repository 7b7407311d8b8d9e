import sys
sys.path.insert(0, "/root/repo/kobato-eyes_b200")
import torch
from kobato_b200 import ops
bank = ops.synth_images_device(0, 8192, 512, 512, 1, n_set=8192)
ops.phash_dhash_batch(bank[:64]); ops.phash_dhash_batch(bank); torch.cuda.synchronize()
