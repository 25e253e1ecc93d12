"""Three parties, files in between: a developer compiles a program, a key owner (client) generates keys,
encrypts and later decrypts, an untrusted server with a B200 evaluates.  Every object that changes hands goes
through save() / load() (the reference's protobuf file format, eva_b200/serialization.py).

    python examples/serialization.py [workdir]      # needs a GPU
"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from eva import EvaProgram, Input, Output, evaluate, load, save  # noqa: E402
from eva.ckks import CKKSCompiler  # noqa: E402
from eva.metric import valuation_mse  # noqa: E402
from eva.seal import generate_keys  # noqa: E402

VEC = 8


class Exchange:
    """the shared directory the parties drop files into"""

    def __init__(self, root):
        self.root = root

    def put(self, name, obj):
        save(obj, os.path.join(self.root, name))

    def get(self, name):
        return load(os.path.join(self.root, name))


def developer(box):
    """writes program.eva, program.evaparams, program.evasignature"""
    quadratic = EvaProgram('Quadratic', vec_size=VEC)
    with quadratic:
        t = Input('t')
        Output('height', -4.9 * t ** 2 + 12 * t + 1.5)
    quadratic.set_input_scales(20)
    quadratic.set_output_ranges(20)
    compiled, params, signature = CKKSCompiler().compile(quadratic)
    for name, obj in (('program.eva', compiled), ('program.evaparams', params), ('program.evasignature', signature)):
        box.put(name, obj)


def client_setup(box, samples):
    """keys from the published parameters; the encrypted samples and the PUBLIC context go to the server"""
    public_ctx, secret_ctx = generate_keys(box.get('program.evaparams'))
    box.put('client.sealsecret', secret_ctx)          # stays with the client
    box.put('server.sealpublic', public_ctx)
    signature = box.get('program.evasignature')
    box.put('request.sealvals', load(os.path.join(box.root, 'server.sealpublic')).encrypt(samples, signature))


def server(box):
    """sees only ciphertexts, evaluation keys and the program"""
    ctx = box.get('server.sealpublic')
    box.put('response.sealvals', ctx.execute(box.get('program.eva'), box.get('request.sealvals')))


def client_finish(box):
    return box.get('client.sealsecret').decrypt(box.get('response.sealvals'), box.get('program.evasignature'))


def main(workdir):
    box = Exchange(workdir)
    samples = {'t': [0.25 * i for i in range(VEC)]}
    developer(box)
    client_setup(box, samples)
    server(box)
    heights = client_finish(box)
    expected = evaluate(box.get('program.eva'), samples)
    mse = valuation_mse(heights, expected)
    print('expected ', [round(v, 3) for v in expected['height']])
    print('decrypted', [round(v, 3) for v in heights['height']])
    print('MSE', mse)
    assert mse < 0.01
    return mse


if __name__ == '__main__':
    if len(sys.argv) > 1:
        main(sys.argv[1])
    else:
        with tempfile.TemporaryDirectory() as scratch:
            main(scratch)
