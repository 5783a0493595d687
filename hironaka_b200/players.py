"""Parameter-free host / agent modules, same behaviour as
hironaka/trainer/player_modules/modules.py:15-79 (used by Trainer.evaluate_rho and by tests)."""
from __future__ import annotations

import torch
from torch import nn


class _Player(nn.Module):
    def __init__(self, dimension: int, max_num_points: int, device: torch.device):
        super().__init__()
        self.dimension = dimension
        self.max_num_points = max_num_points
        self.device = device


class ChooseFirstAgentModule(_Player):
    def forward(self, x):
        return nn.functional.one_hot(x["coords"].argmax(1), num_classes=self.dimension).type(torch.float32)


class ChooseLastAgentModule(_Player):
    def forward(self, x):
        c = x["coords"].type(torch.float32)
        aug = c + torch.arange(self.dimension, dtype=torch.float32, device=c.device) * 1e-4
        return nn.functional.one_hot(aug.argmax(1), num_classes=self.dimension).type(torch.float32)


class RandomAgentModule(_Player):
    def forward(self, x):
        c = x["coords"]
        r = torch.rand((c.shape[0], self.dimension), device=c.device) * c
        return nn.functional.one_hot(r.argmax(1), num_classes=self.dimension).type(torch.float32)


class RandomHostModule(_Player):
    def __init__(self, dimension, max_num_points, device):
        super().__init__(dimension, max_num_points, device)
        self.output_dim = 2 ** dimension - dimension - 1

    def forward(self, x):
        r = torch.randint(self.output_dim, (x.shape[0],), device=x.device)
        return nn.functional.one_hot(r.long(), num_classes=self.output_dim).type(torch.float32)


class AllCoordHostModule(_Player):
    def __init__(self, dimension, max_num_points, device):
        super().__init__(dimension, max_num_points, device)
        self.output_dim = 2 ** dimension - dimension - 1

    def forward(self, x):
        r = torch.zeros((x.shape[0], self.output_dim), device=x.device, dtype=torch.float32)
        r[:, -1] = 1.0
        return r
