"""The client / server flow of the reference's examples/serialization.py with every object going
through files (save / load), executed by the B200 backend.  Run on a GPU box:
    python examples/serialization.py [workdir]
"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from eva import EvaProgram, Input, Output, evaluate, save, load  # noqa: E402
from eva.ckks import CKKSCompiler  # noqa: E402
from eva.seal import generate_keys  # noqa: E402
from eva.metric import valuation_mse  # noqa: E402


def main(workdir):
    path = lambda n: os.path.join(workdir, n)
    # ---- compile time
    poly = EvaProgram('Polynomial', vec_size=8)
    with poly:
        x = Input('x')
        Output('y', 3 * x ** 2 + 5 * x - 2)
    poly.set_output_ranges(20)
    poly.set_input_scales(20)
    poly, params, signature = CKKSCompiler().compile(poly)
    save(poly, path('poly.eva'))
    save(params, path('poly.evaparams'))
    save(signature, path('poly.evasignature'))
    # ---- key generation time
    public_ctx, secret_ctx = generate_keys(load(path('poly.evaparams')))
    save(public_ctx, path('poly.sealpublic'))
    save(secret_ctx, path('poly.sealsecret'))
    # ---- runtime on the client
    signature = load(path('poly.evasignature'))
    public_ctx = load(path('poly.sealpublic'))
    inputs = {'x': [i for i in range(signature.vec_size)]}
    save(public_ctx.encrypt(inputs, signature), path('poly_inputs.sealvals'))
    # ---- runtime on the server
    poly = load(path('poly.eva'))
    public_ctx = load(path('poly.sealpublic'))
    enc_outputs = public_ctx.execute(poly, load(path('poly_inputs.sealvals')))
    save(enc_outputs, path('poly_outputs.sealvals'))
    # ---- back on the client
    secret_ctx = load(path('poly.sealsecret'))
    outputs = secret_ctx.decrypt(load(path('poly_outputs.sealvals')), signature)
    reference = evaluate(poly, inputs)
    mse = valuation_mse(outputs, reference)
    print('Expected', reference)
    print('Got', outputs)
    print('MSE', mse)
    assert mse < 0.01
    return mse


if __name__ == '__main__':
    if len(sys.argv) > 1:
        main(sys.argv[1])
    else:
        with tempfile.TemporaryDirectory() as d:
            main(d)
