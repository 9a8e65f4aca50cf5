"""Run-time switches of espm_b200 (module-level so that the estimator keeps the reference's signature).

distributed : "auto" | True | False
    Shard the image rows over the ranks of the default torch.distributed process group.  "auto" shards
    whenever a process group with more than one rank is initialised.
"""
distributed = "auto"
