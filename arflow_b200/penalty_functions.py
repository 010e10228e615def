"""Drop-in for losses/penalty_functions.py of deu439/ARFlow (scalar penalty functions and their lookup)."""
import torch


def abs_robust_loss(diff, eps=0.01, q=0.4):
    """(|d| + eps)^q — penalty_functions.py:3-4."""
    return torch.pow((torch.abs(diff) + eps), q)


def charbonnier(x_sq, eps=0.001):
    """sqrt(x^2 + eps^2), applied to an already squared argument — :6-7."""
    return torch.sqrt(x_sq + eps ** 2)


def charbonnier_prime(x_sq, eps=0.001):
    """:9-10"""
    return 1 / (2 * torch.sqrt(x_sq + eps ** 2))


def identity(x):
    return x


def identity_prime(x):
    return torch.ones_like(x)


_TABLE = {'identity': (identity, identity_prime), 'charbonnier': (charbonnier, charbonnier_prime),
          'abs_robust_loss': (abs_robust_loss, None)}


def get_penalty(name, derivative=False):
    """:18-28 — unknown names return None like the reference's fall-through."""
    if name not in _TABLE:
        return None
    fn, prime = _TABLE[name]
    if not derivative:
        return fn
    if prime is None:
        raise NotImplementedError("derivative not implemented ofr abs_robust_loss penalty!")
    return prime
