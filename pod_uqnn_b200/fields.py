"""Registry of the analytic snapshot fields u(X, t, mu) the CUDA path evaluates.

The reference passes an arbitrary Python callable to numba (podnnmodel.py:99); a CUDA
drop-in can only serve a closed registry.  A field is named by a `Field` object, by its
string name, or by any callable tagged with a ``pb_field`` attribute
(``u.pb_field = "ackley"``).  Anything else raises NotImplementedError -- there is no CPU
fallback.
"""
from ._lib import FIELD_IDS


class Field:
    """Handle of a registered field; `cite` is the reference definition it reproduces."""

    def __init__(self, name, dim, has_t, cite):
        self.name, self.dim, self.has_t, self.cite = name, dim, has_t, cite
        self.field_id = FIELD_IDS[name]
        self.pb_field = name

    def __call__(self, X, t, mu):
        raise NotImplementedError(
            f"field '{self.name}' is evaluated on the GPU by loop_u/loop_u_t; it has no host implementation")

    def __repr__(self):
        return f"Field({self.name!r})"


u_shekel = Field("shekel", 1, False, "experiments/1d_shekel/hyperparams.py:55-67")
u_ackley = Field("ackley", 2, False, "experiments/2d_ackley/hyperparams.py:51-57")
u_burgers = Field("burgers", 1, True, "experiments/1dt_burger/hyperparams.py:45-55")
REGISTRY = {f.name: f for f in (u_shekel, u_ackley, u_burgers)}


def resolve(u):
    if isinstance(u, Field):
        return u
    name = u if isinstance(u, str) else getattr(u, "pb_field", None)
    if isinstance(name, str) and name in REGISTRY:
        return REGISTRY[name]
    raise NotImplementedError(
        f"snapshot field {u!r} is not registered for the CUDA path (known: {sorted(REGISTRY)}); "
        "tag a reference field with u.pb_field = '<name>' or pass the name.  No CPU fallback.")
