"""Mirror of RAGraph_node/ragraph_utils/TaskDecoder.py:3-17 (dense MLP: stays a torch module)."""
import torch.nn as nn


class TaskDecoder(nn.Module):
    def __init__(self, input_dim, hiddden_dim, output_dim):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, hiddden_dim)
        self.act = nn.LeakyReLU()
        self.fc2 = nn.Linear(hiddden_dim, output_dim)

    def reset_parameters(self):
        self.fc1.reset_parameters()
        self.fc2.reset_parameters()

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))
