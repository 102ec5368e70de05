"""Mirror of the reference's ``help_classes`` module (help_classes.py:1-74): the plain record types the SfM tables hold.
Same names, constructor arguments and attributes (including the reference's initial ``observations_index = [0]``)."""
import numpy as np


class CameraPose:
    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else R
        self.t = np.array([0.0, 0.0, 0.0]) if t is None else t

    def __str__(self):
        return "R: " + str(self.R) + " t: " + str(self.t)

    def GetCameraMatrix(self):
        C = np.zeros((3, 4), dtype='double')
        C[:, -1] = self.t
        C[:3, :3] = self.R
        return C


class Point_3D:
    def __init__(self, point):
        self.point = point
        self.observations_index = np.array([0], dtype='int')

    def __str__(self):
        return "3D point: " + str(self.point) + " Observation index: " + str(self.observations_index)


class Observation:
    def __init__(self, image_coordinates, view_index, point_3D_index, color):
        self.image_coordinates = image_coordinates
        self.view_index = view_index
        self.point_3D_index = point_3D_index
        self.color = color

    def __str__(self):
        return ("OBSERVATION: Image coords: " + str(self.image_coordinates) + " View index: " + str(self.view_index) +
                " 3D point index " + str(self.point_3D_index) + " Color: " + str(self.color))


class View:
    def __init__(self, image, camera_pose):
        self.image = image
        self.camera_pose = camera_pose
        self.observations_index = np.array([0], dtype='int')

    def getWorldPosition(self):
        return -1.0 * (self.camera_pose.R.T @ self.camera_pose.t)

    def __str__(self):
        return ("VIEW: Image index: " + str(self.image) + " Camera pose: " + str(self.camera_pose) +
                " Observations table " + str(self.observations_index))
