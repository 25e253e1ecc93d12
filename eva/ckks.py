from eva_b200.ckks import *  # noqa: F401,F403
from eva_b200.ckks import CKKSCompiler, CKKSEncodingInfo, CKKSParameters, CKKSSignature  # noqa: F401
