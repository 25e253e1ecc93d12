"""`eva.seal` of the reference, served by the B200 backend (same call signatures)."""
from eva_b200.b200 import generate_keys  # noqa: F401
from eva_b200.b200 import B200Public as SEALPublic, B200Secret as SEALSecret, B200Valuation as SEALValuation  # noqa: F401
