"""Models package (reference: neural_jump_ode/models/__init__.py:3)."""

from .jump_ode import NeuralJumpODE, JumpNN, ODEFunc, OutputNN, nj_ode_loss

__all__ = ["NeuralJumpODE", "JumpNN", "ODEFunc", "OutputNN", "nj_ode_loss"]
