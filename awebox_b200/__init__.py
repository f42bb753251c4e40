"""awebox_b200 — B200-native evaluator for awebox's direct-collocation constraint residual g(w,p)
and its sparse Jacobian."""
__version__ = '0.1.0'
