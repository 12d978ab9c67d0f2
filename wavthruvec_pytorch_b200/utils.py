"""Constructor helpers with the reference's semantics (vec2wav/utils.py:23-26,35-36)."""
from __future__ import annotations


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    """'same' padding of a dilated odd-size kernel: (k*d - d)/2  (vec2wav/utils.py:35-36)."""
    return (kernel_size * dilation - dilation) // 2


def init_weights(m, mean: float = 0.0, std: float = 0.01) -> None:
    """N(mean, std) draw into ``m.weight`` of every ``*Conv*`` module (vec2wav/utils.py:23-26).

    On a weight-normed module ``m.weight`` is the *derived* tensor, not a parameter, so in
    the reference this call only advances the RNG; ``weight_g/weight_v`` keep their default
    init.  It is reproduced call-for-call because bit-identical construction under a seed
    (the golden fixtures depend on it) needs the same RNG consumption.
    """
    if "Conv" in type(m).__name__:
        m.weight.data.normal_(mean, std)
