"""The benchmark / example programs written against the user-facing DSL (our own
formulation of reference examples/image_processing.py:12-100, README.md:93-105 and the
wide DAG of BASELINE config 5)."""
from eva import EvaProgram, Input, Output

SOBEL_FILTER = [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]


def _conv_xy(image, width, filt):
    ix = iy = None
    for i in range(3):
        for j in range(3):
            rot = image << (i * width + j)
            h, v = rot * filt[i][j], rot * filt[j][i]
            ix = h if ix is None else ix + h
            iy = v if iy is None else iy + v
    return ix, iy


def _conv(image, width, filt):
    acc = None
    for i in range(3):
        for j in range(3):
            part = (image << (i * width + j)) * filt[i][j]
            acc = part if acc is None else acc + part
    return acc


def sobel(h=64, w=64):
    prog = EvaProgram('sobel', vec_size=h * w)
    with prog:
        image = Input('image')
        ix, iy = _conv_xy(image, w, SOBEL_FILTER)
        d = ix ** 2 + iy ** 2
        d2 = d * d
        d3 = d2 * d
        Output('image', d * 2.2137874823876622 + d2 * -1.0984324107372518 + d3 * 0.17254603006834726)
    prog.set_input_scales(25)
    prog.set_output_ranges(10)
    return prog


def harris(h=64, w=64):
    prog = EvaProgram('harris', vec_size=h * w)
    with prog:
        image = Input('image')
        pool = [[1, 1, 1]] * 3
        ix, iy = _conv_xy(image, w, SOBEL_FILTER)
        ixx, iyy, ixy = ix ** 2, iy ** 2, ix * iy
        sxx, syy, sxy = _conv(ixx, w, pool), _conv(iyy, w, pool), _conv(ixy, w, pool)
        det = sxx * syy - sxy * sxy
        trace = sxx + syy
        Output('image', det - trace ** 2 * 0.04)
    prog.set_input_scales(30)
    prog.set_output_ranges(20)
    return prog


def polynomial():
    prog = EvaProgram('Polynomial', vec_size=1024)
    with prog:
        x = Input('x')
        Output('y', 3 * x ** 2 + 5 * x - 2)
    prog.set_output_ranges(30)
    prog.set_input_scales(30)
    return prog


def wide(nprod):
    prog = EvaProgram('wide%d' % nprod, vec_size=8192)
    with prog:
        x, y = Input('x'), Input('y')
        terms = [(x << (i % 64)) * (y << ((i // 64) % 64)) for i in range(nprod)]
        while len(terms) > 1:
            terms = [terms[i] + terms[i + 1] for i in range(0, len(terms), 2)]
        Output('z', terms[0])
    prog.set_input_scales(40)
    prog.set_output_ranges(30)
    return prog
