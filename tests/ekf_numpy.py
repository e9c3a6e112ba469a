"""Independent NumPy float64 restatement of the reference EKF, written in the
reference's own matrix form (F, Gxm, Gx = Gxm F, K = S Gx^T (Gx S Gx^T + R)^-1,
(I - K Gx) S) -- src/aruco_slam.cpp:21-74 and :88-263.  Test helper only."""
import numpy as np

PI = 3.14159265358979323846


def norm_angle(a):
    if a >= PI:
        a -= 2 * PI
    if a < -PI:
        a += 2 * PI
    return a


def _push_heap(h, v, less):
    """libstdc++ std::push_heap (bits/stl_heap.h __push_heap): sift the new last element up"""
    h.append(v)
    hole = len(h) - 1
    parent = (hole - 1) // 2
    while hole > 0 and less(h[parent], v):
        h[hole] = h[parent]
        hole = parent
        parent = (hole - 1) // 2
    h[hole] = v


def _pop_heap(h, less):
    """libstdc++ std::pop_heap + pop_back (__adjust_heap: sift the hole down to a leaf, then the old last element up)"""
    top = h[0]
    v = h.pop()
    n = len(h)
    if n == 0:
        return top
    hole = child = 0
    while child < (n - 1) // 2:
        child = 2 * (child + 1)
        if less(h[child], h[child - 1]):
            child -= 1
        h[hole] = h[child]
        hole = child
    if n % 2 == 0 and child == (n - 2) // 2:
        child = 2 * (child + 1)
        h[hole] = h[child - 1]
        hole = child - 1
    parent = (hole - 1) // 2
    while hole > 0 and less(h[parent], v):
        h[hole] = h[parent]
        hole = parent
        parent = (hole - 1) // 2
    h[hole] = v
    return top


class NumpyEkf:
    def __init__(self, Q_k=0.01, kl=0.05, kr=0.05, b=0.09):
        self.Q_k, self.kl, self.kr, self.b = Q_k, kl, kr, b
        self.mu = np.zeros(3)
        self.sigma = np.zeros((3, 3))
        self.id_map = {}
        self.last = []     # last_observed_marker_: (id, last_observation_ or None) in the order they were processed

    def predict(self, wl, wr, dt):
        dsl, dsr = self.kl * dt * wl, self.kr * dt * wr
        dth = (dsr - dsl) / (2 * self.b)
        ds = 0.5 * (dsr + dsl)
        th = self.mu[2] + 0.5 * dth
        c, s = np.cos(th), np.sin(th)
        self.mu[0] += ds * c
        self.mu[1] += ds * s
        self.mu[2] = norm_angle(self.mu[2] + dth)
        N = len(self.mu)
        Hx = np.eye(N)
        Hx[:3, :3] = [[1, 0, -ds * s], [0, 1, ds * c], [0, 0, 1]]
        wkh = (0.5 * self.kl * dt) * np.array([[c, c], [s, s], [1 / self.b, -1 / self.b]])
        su = np.diag([self.Q_k * abs(wl), self.Q_k * abs(wr)])
        F = np.zeros((N, 3))
        F[:3, :3] = np.eye(3)
        self.sigma = Hx @ self.sigma @ Hx.T + F @ (wkh @ su @ wkh.T) @ F.T

    def update(self, obs):
        """obs: list of (id, x, y, theta, R 3x3) in detection order; processed in the order the reference's
        std::priority_queue (operator< = index greater, aruco_slam.h:85-88) pops them under libstdc++"""
        heap = []
        less = lambda a, b: a[0] > b[0]
        for seq, (aid, x, y, th, R) in enumerate(obs):
            _push_heap(heap, (self.id_map.get(aid, -1), seq, aid, x, y, th, np.asarray(R, float).reshape(3, 3)), less)
        items = [_pop_heap(heap, less) for _ in range(len(obs))]
        mu = self.mu.copy()
        new_last = []
        for idx, _, aid, ox, oy, oth, Rk in items:
            lastobs = None
            if idx >= 0:
                N = len(self.mu)
                F = np.zeros((6, N))
                F[:3, :3] = np.eye(3)
                F[3:, 3 + 3 * idx:6 + 3 * idx] = np.eye(3)
                mx, my, mth = mu[3 + 3 * idx:6 + 3 * idx]
                x, y, th = mu[:3]
                s, c = np.sin(th), np.cos(th)
                gdx, gdy, gdt = mx - x, my - y, norm_angle(mth - th)
                zhat = np.array([gdx * c + gdy * s, -gdx * s + gdy * c, gdt])
                z = np.array([ox, oy, oth])
                ze = z - zhat
                ze[2] = norm_angle(ze[2])
                Gxm = np.array([[-c, -s, -gdx * s + gdy * c, c, s, 0],
                                [s, -c, -gdx * c - gdy * s, -s, c, 0],
                                [0, 0, -1, 0, 0, 1]])
                Gx = Gxm @ F
                K = self.sigma @ Gx.T @ np.linalg.inv(Gx @ self.sigma @ Gx.T + Rk)
                prev = next((lo for (i, lo) in self.last if i == aid), None)          # std::find: the first entry with this id
                stationary = prev is not None and np.linalg.norm(prev - z) < 0.01
                if not stationary:
                    lastobs = z
                    self.mu = self.mu + K @ ze
                    self.sigma = (np.eye(N) - K @ Gx) @ self.sigma
            else:
                sinth, costh = float(np.float32(np.sin(mu[2]))), float(np.float32(np.cos(mu[2])))
                N = len(self.mu)
                map_x = mu[0] + costh * ox - sinth * oy
                map_y = mu[1] + sinth * ox + costh * oy
                map_th = norm_angle(mu[2] + oth)
                dx, dy = map_x - mu[0], map_y - mu[1]
                Gsk = np.array([[-costh, -sinth, -sinth * dx + costh * dy],
                                [sinth, -costh, -dx * costh - dy * sinth], [0, 0, -1]])
                Gmi = np.array([[costh, sinth, 0], [-sinth, costh, 0], [0, 0, 1]])
                Smm = Gmi @ (Gsk @ self.sigma[:3, :3] @ Gsk.T + Rk).T @ Gmi.T
                Smx = -Gmi @ Gsk @ self.sigma[:3, :]
                ns = np.zeros((N + 3, N + 3))
                ns[:N, :N] = self.sigma
                ns[:N, N:] = Smx.T
                ns[N:, :N] = Smx
                ns[N:, N:] = Smm
                self.sigma = ns
                self.mu = np.concatenate([self.mu, [map_x, map_y, map_th]])
                self.id_map.setdefault(aid, (len(self.mu) - 3) // 3 - 1)   # std::map::insert keeps the first
            new_last.append((aid, lastobs))
        self.last = new_last
