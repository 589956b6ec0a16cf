"""Task decoder head of RAGraph.forward: the dense MLP that turns the fused hidden embedding into class logits.

Interface parity with RAGraph_node/ragraph_utils/TaskDecoder.py:3-17 (constructor arguments, ``reset_parameters``, and the
``fc1`` / ``fc2`` parameter names, so a reference ``state_dict`` loads unchanged).  It is a two-layer perceptron on
[Q, d] activations -- a plain library GEMM, outside the sparse / retrieval hot path -- so it stays a torch module; nothing
here calls the CUDA library.
"""
from collections import OrderedDict

import torch
from torch import nn


class TaskDecoder(nn.Module):
    """logits = W2 . leaky_relu(W1 . h + b1, 0.01) + b2"""

    NEGATIVE_SLOPE = 0.01       # nn.LeakyReLU default, what the reference instantiates

    def __init__(self, input_dim: int, hiddden_dim: int, output_dim: int) -> None:
        super().__init__()
        layers = OrderedDict(fc1=nn.Linear(input_dim, hiddden_dim), act=nn.LeakyReLU(self.NEGATIVE_SLOPE),
                             fc2=nn.Linear(hiddden_dim, output_dim))
        for name, layer in layers.items():          # registered under the reference's attribute names
            self.add_module(name, layer)

    def reset_parameters(self) -> None:
        for layer in (self.fc1, self.fc2):
            layer.reset_parameters()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(self.act(self.fc1(x)))
