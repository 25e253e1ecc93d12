"""CKKS compiler front-end types (reference python/eva/ckks/__init__.py)."""
from ._eva_b200._ckks import CKKSEncodingInfo, CKKSParameters, CKKSSignature  # noqa: F401
