"""tests/golden/flocking_models.npz: the ten state dicts of data/models/experiment_Flocking-seed_*.pth of the reference
(three GATConv layers 7 -> 8 -> 8 -> 8 plus two Linear layers; the class that produced them is not in the repository, so
they serve as realistic LAYER weights for the generic GAT layer tests, not as a network with known activations).
Run in the build container (reads /root/reference):  python tests/golden/make_flocking_models.py"""
import os

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    out = {}
    for seed in range(10):
        sd = torch.load(f"{REF}/data/models/experiment_Flocking-seed_{seed}.pth", map_location="cpu")
        for k, v in sd.items():
            out[f"{seed}/{k}"] = v.numpy()
    np.savez_compressed(f"{HERE}/flocking_models.npz", **out)
    print(len(out), "tensors ->", f"{HERE}/flocking_models.npz", os.path.getsize(f"{HERE}/flocking_models.npz"), "bytes")
