"""Constants of the fit loop (values of espm/conf.py:55-59)."""
log_shift = 1e-14
dicotomy_tol = 1e-5
sigmaL = 8
maxit_dichotomy = 100
