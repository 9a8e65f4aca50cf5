"""Run-time switches of espm_b200 (module-level so that the estimator keeps the reference's signature).

distributed : "auto" | True | False
    Shard the image rows over the ranks of the default torch.distributed process group.  "auto" shards
    whenever a process group with more than one rank is initialised.
device_init : bool
    When neither W nor H is passed to ``fit_transform``, run the randomized SVD of scikit-learn's NNDSVD
    initialisation on the device (init_device.py) instead of on the host.  Unsharded fits with n < p only.
"""
distributed = "auto"
device_init = True
