"""Drop-in import name of the reference package: `from OpenGoddard.optimize import ...`
resolves to the B200-native engine in `opengoddard_b200`."""
