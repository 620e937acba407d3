"""Opponent pool (reference: src/selfplay/opponent_pool.py:5-19): at most `max_size` frozen policies, the oldest replaced
first, one drawn uniformly at random.  Host-side bookkeeping only; one opponent serves all envs for a whole iteration
(src/train.py:106-114).  Kept as a fixed ring with a write cursor (entries are whole policies holding device weights: nothing
is shifted or re-allocated when the pool is full); `pool` lists the entries oldest first, and a draw consumes the global
`random` stream exactly as `random.choice` over the reference's deque does, so seeded runs pick the same opponents."""
import random


class OpponentPool:
    def __init__(self, max_size=5):
        self.max_size = max_size
        self._ring = []           # at most max_size entries
        self._oldest = 0          # position of the oldest entry once the ring is full

    @property
    def pool(self):
        return self._ring[self._oldest:] + self._ring[:self._oldest]

    def size(self):
        return len(self._ring)

    def add_opponent(self, opponent):
        if self.max_size is not None and len(self._ring) >= self.max_size:
            self._ring[self._oldest] = opponent
            self._oldest = (self._oldest + 1) % len(self._ring)
        else:
            self._ring.append(opponent)

    def get_random_opponent(self):
        count = len(self._ring)
        if count == 0:
            return None
        return self._ring[(self._oldest + random.randrange(count)) % count]
