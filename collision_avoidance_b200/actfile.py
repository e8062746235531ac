"""Reader / writer of ALAN action-set files (``*.act``).

Format (written by the reference with ``f.write(str(actions))``,
collision_avoidance/ALAN/Train_ALAN_action_space.py:154-156): one line holding the Python
literal of a list of ``(x, y)`` tuples; the first action is always ``(1, 0)``.
"""
from __future__ import annotations

import ast
from typing import List, Sequence, Tuple

Action = Tuple[float, float]


def loads(text: str) -> List[Action]:
    data = ast.literal_eval(text.strip())
    if not isinstance(data, (list, tuple)) or not data:
        raise ValueError("an .act file holds a non-empty list of (x, y) tuples")
    out = []
    for item in data:
        if not isinstance(item, (list, tuple)) or len(item) != 2:
            raise ValueError(f"bad action entry {item!r}")
        out.append((float(item[0]), float(item[1])))
    return out


def dumps(actions: Sequence[Action]) -> str:
    # str(list of tuples) exactly like the reference; ints stay ints, e.g. "(1, 0)"
    return str([tuple(a) for a in actions])


def load(path: str) -> List[Action]:
    with open(path) as f:
        return loads(f.read())


def save(path: str, actions: Sequence[Action]) -> None:
    with open(path, "w+") as f:
        f.write(dumps(actions))
