"""Host-side API logic that needs no GPU: constructor forms, argument errors."""
import pytest


def test_correlation_constructor_forms():
    """Both reference constructors (correlation_package/correlation.py:48, correlation_native.py:7)."""
    from arflow_b200.correlation import Correlation
    c = Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)   # pwclite.py:124-126
    assert (c.pad_size, c.kernel_size, c.max_displacement, c.stride1, c.stride2) == (4, 1, 4, 1, 1)
    n = Correlation(3)                       # native form: the positional argument is max_displacement
    assert (n.pad_size, n.kernel_size, n.max_displacement, n.stride1, n.stride2) == (3, 1, 3, 1, 1)
    n = Correlation()                        # native default
    assert (n.pad_size, n.max_displacement) == (4, 4)
    n = Correlation(max_displacement=4, kernel_size=1, stride1=1, stride2=1, corr_multiply=1)   # native swallows the rest
    assert (n.pad_size, n.kernel_size, n.stride2) == (4, 1, 1)
    p = Correlation(4, 1, 4)                 # package form, positional: stride2 keeps the package default 2
    assert (p.pad_size, p.kernel_size, p.max_displacement, p.stride1, p.stride2) == (4, 1, 4, 1, 2)
    p = Correlation(pad_size=2)              # package defaults: kernel_size 0, max_displacement 0, stride2 2
    assert (p.kernel_size, p.max_displacement, p.stride2) == (0, 0, 2)
    with pytest.raises(TypeError):
        Correlation(bogus=1)
    with pytest.raises(TypeError):
        Correlation(4, 1, pad_size=4)
