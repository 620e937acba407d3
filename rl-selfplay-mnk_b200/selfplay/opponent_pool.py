"""Opponent pool (reference: src/selfplay/opponent_pool.py:5-19): a bounded FIFO of frozen
policies with uniform random choice.  Host-side bookkeeping only; one opponent serves all envs
for a whole iteration (src/train.py:106-114)."""
import random
from collections import deque


class OpponentPool:
    def __init__(self, max_size=5):
        self.max_size = max_size
        self.pool = deque(maxlen=max_size)      # oldest entry is evicted first

    def add_opponent(self, opponent):
        self.pool.append(opponent)

    def get_random_opponent(self):
        return random.choice(self.pool) if self.pool else None

    def size(self):
        return len(self.pool)
