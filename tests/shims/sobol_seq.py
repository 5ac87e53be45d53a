"""Stand-in for sobol_seq.i4_sobol_generate (test infrastructure, see README.md): the first n points of an unscrambled
Sobol sequence after the origin."""
from scipy.stats import qmc


def i4_sobol_generate(dim_num, n, skip=1):
    s = qmc.Sobol(d=int(dim_num), scramble=False)
    if skip:
        s.fast_forward(int(skip))
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return s.random(int(n))
