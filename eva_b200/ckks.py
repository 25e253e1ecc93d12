"""CKKS compiler front-end (reference python/eva/ckks/__init__.py)."""
from ._eva_b200._ckks import CKKSCompiler, CKKSEncodingInfo, CKKSParameters, CKKSSignature  # noqa: F401
