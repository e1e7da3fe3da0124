from cryovit_b200.host.infer import run_inference  # noqa: F401
