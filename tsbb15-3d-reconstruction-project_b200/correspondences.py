"""Mirror of the reference's ``correspondences`` module (correspondences.py:1-35): the loader of the tracked Dino points
that ``main.py:28-30,113`` feeds into the accelerated path.  Pure host-side data plumbing, kept so that the flat-module
drop-in covers every name ``main.py`` imports; same constructor behaviour (reads ``BAdino2.mat`` from the working
directory, the branch the reference ships active) and the same ``getCorrByIndices(i1, i2)`` -> ``(y1, y2)`` of shape
(N, 2) with the rows removed where either view has no observation (-1)."""
import numpy as np


class Correspondences:
    def __init__(self, path='BAdino2.mat'):
        import scipy.io as sio
        points = sio.loadmat(path)
        self.points = np.asarray(points['newPoints2D'].tolist())[0, :, :, :]       # (views, 2, tracks)

    def getCorrByIndices(self, i1, i2):
        y1 = self.points[i1, :, :].T
        y2 = self.points[i2, :, :].T
        seen = np.logical_and(np.any(y1 != -1, axis=1), np.any(y2 != -1, axis=1))
        return np.array(y1[seen, :]), np.array(y2[seen, :])
