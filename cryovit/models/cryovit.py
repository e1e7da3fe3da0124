from cryovit_b200.host.models import CryoVIT  # noqa: F401
