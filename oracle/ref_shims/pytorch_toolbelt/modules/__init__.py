import torch.nn as nn


def instantiate_activation_block(name, **kwargs):
    table = {"none": nn.Identity, "relu": nn.ReLU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh}
    return table[str(name).lower()]()
