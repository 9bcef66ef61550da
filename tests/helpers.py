"""Shared test helpers (test infrastructure; may import the oracle)."""
import numpy as np
import torch


def state_from_npz(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith(prefix)}


class TinyCNN(torch.nn.Module):
    """same architecture as tests/golden/make_golden.py::TinyCNN (weights come from the fixture)."""

    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(
            torch.nn.Conv2d(3, 8, 3, stride=2, padding=1), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
            torch.nn.Conv2d(8, 16, 3, stride=2, padding=1), torch.nn.BatchNorm2d(16), torch.nn.ReLU(),
            torch.nn.AdaptiveAvgPool2d(2), torch.nn.Flatten())
        self.fc = torch.nn.Linear(64, 2)

    def forward(self, x):
        return self.fc(self.net(x))


def tiny_cnn_from(z, prefix):
    m = TinyCNN()
    m.load_state_dict(state_from_npz(z, prefix))
    return m.eval()
