"""`OpenGoddard.optimize` -- same public names as the reference module
(/root/reference/OpenGoddard/optimize.py), implemented by opengoddard_b200."""
from opengoddard_b200.optimize import Problem, Guess, Condition, Dynamics  # noqa: F401

__all__ = ["Problem", "Guess", "Condition", "Dynamics"]
