"""Oracle: numpy restatement of largesteps/optimize.py AdamUniform (TEST INFRASTRUCTURE)."""
import numpy as np

f32 = np.float32


class AdamUniformOracle:
    """optimize.py:3-41.  Adam whose update is divided by the scalar max(sqrt(m2)) (optimize.py:40)."""

    def __init__(self, shape, lr=0.1, betas=(0.9, 0.999)):
        self.lr = lr
        self.b1, self.b2 = betas
        self.step_count = 0
        self.g1 = np.zeros(shape, dtype=f32)      # optimize.py:27
        self.g2 = np.zeros(shape, dtype=f32)      # optimize.py:28

    def step(self, p, grad):
        b1, b2 = self.b1, self.b2
        self.step_count += 1                                              # optimize.py:32
        grad = np.asarray(grad, dtype=f32)
        self.g1 = (self.g1 * f32(b1) + f32(1 - b1) * grad).astype(f32)           # optimize.py:35
        self.g2 = (self.g2 * f32(b2) + f32(1 - b2) * (grad * grad)).astype(f32)  # optimize.py:36
        m1 = (self.g1 / f32(1 - b1 ** self.step_count)).astype(f32)              # optimize.py:37
        m2 = (self.g2 / f32(1 - b2 ** self.step_count)).astype(f32)              # optimize.py:38
        gr = (m1 / (f32(1e-8) + np.sqrt(m2).max())).astype(f32)                  # optimize.py:40
        return (np.asarray(p, dtype=f32) - f32(self.lr) * gr).astype(f32)        # optimize.py:41
