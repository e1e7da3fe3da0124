from cryovit_b200.host.datasets import collate_fn  # noqa: F401
