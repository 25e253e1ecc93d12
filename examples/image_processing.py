"""Sobel edge filter and Harris corner detector on an encrypted 64x64 image, executed on a
B200 through the EVA API (the workload of the reference's examples/image_processing.py;
run from the repository root:  python examples/image_processing.py [image.png]).

The script is written against `eva` exactly like a reference user script: only the
backend behind eva.seal.generate_keys / encrypt / execute / decrypt is different."""
import math
import sys
import os

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from eva import evaluate                       # noqa: E402
from eva.ckks import CKKSCompiler              # noqa: E402
from eva.metric import valuation_mse           # noqa: E402
from eva.seal import generate_keys             # noqa: E402
from tests_programs import harris, sobel       # noqa: E402

h = w = 64


def read_input_image(path=None):
    if path:
        from PIL import Image
        image = Image.open(path).convert('L').resize((w, h))
        return {'image': [x / 255.0 for x in list(image.getdata())]}
    # smooth synthetic test pattern
    return {'image': [0.5 + 0.25 * math.sin(0.1 * (k % w)) * math.cos(0.07 * (k // w)) for k in range(h * w)]}


if __name__ == "__main__":
    inputs = read_input_image(sys.argv[1] if len(sys.argv) > 1 else None)
    for prog in (sobel(h, w), harris(h, w)):
        print('Compiling', prog.name)
        compiled, params, signature = CKKSCompiler().compile(prog)
        print('  N =', params.poly_modulus_degree, 'prime_bits =', list(params.prime_bits), 'rotations =', sorted(params.rotations))
        public_ctx, secret_ctx = generate_keys(params)
        enc_inputs = public_ctx.encrypt(inputs, signature)
        enc_outputs = public_ctx.execute(compiled, enc_inputs)     # on the GPU
        outputs = secret_ctx.decrypt(enc_outputs, signature)
        reference = evaluate(compiled, inputs)
        print('  MSE vs plaintext evaluation:', valuation_mse(outputs, reference))
