"""Drop-in alias: ``import eva`` resolves to the B200 backend, so scripts written
against the reference package (examples/image_processing.py, tests/*.py) run
unchanged:  eva.ckks.CKKSCompiler, eva.seal.generate_keys, eva.metric, eva.std."""
from eva_b200 import *  # noqa: F401,F403
from eva_b200 import (EvaProgram, Expr, Input, Op, Output, Program, Term, Type, evaluate, load, py_to_eva,  # noqa: F401
                      save, set_num_threads)
