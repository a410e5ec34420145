from .mlp import MLP

__all__ = ["MLP"]
