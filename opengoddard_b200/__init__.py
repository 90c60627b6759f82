"""OpenGoddard-B200: a B200-native (sm_100a) batched pseudospectral-collocation
constraint / forward-difference-Jacobian engine behind the OpenGoddard API.

    from OpenGoddard.optimize import Problem, Guess, Condition, Dynamics   # drop-in
"""
__version__ = "0.1.0"
