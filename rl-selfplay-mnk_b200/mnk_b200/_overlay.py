"""Overlay support for the drop-in packages ``env``, ``selfplay`` and ``alg``.

The reference's ``src/env``, ``src/selfplay`` and ``src/alg`` are namespace packages (no
``__init__.py``); the drop-in packages of the same names are regular packages and therefore win the
import regardless of ``sys.path`` order.  A regular package hides every sibling module it does not
define itself, though -- ``alg.ppo``, ``alg.architectures`` or ``selfplay.opponent_pool`` would stop
resolving.  ``extend`` appends the same-named directories found further along ``sys.path`` to the
package's ``__path__``: modules the drop-in defines (``alg.rollout_buffer``, ``env.torch_vector_mnk_env``,
...) come from here, everything else (PPO, networks, schedulers) from the reference, unmodified
(INTEGRATION.md, section A).
"""
import os
import sys


def extend(name: str, path: list) -> None:
    own = {os.path.abspath(p) for p in path}
    for entry in sys.path:
        cand = os.path.abspath(os.path.join(entry or ".", *name.split(".")))
        if cand not in own and os.path.isdir(cand):
            path.append(cand)
            own.add(cand)
