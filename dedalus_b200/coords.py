"""Coordinates (reference: dedalus/core/coords.py:20-190, Cartesian subset)."""
import numpy as np


class Coordinate:
    dim = 1

    def __init__(self, name, cs=None):
        self.name = name
        self.cs = cs
        self.coords = (self,)

    def __repr__(self):
        return f"<Coordinate {self.name}>"

    def check_bounds(self, bounds):
        if bounds[0] >= bounds[1]:
            raise ValueError("Bounds must be increasing.")


class CartesianCoordinates:
    """Cartesian coordinate system (reference coords.py:139-190)."""

    def __init__(self, *names):
        if len(set(names)) < len(names):
            raise ValueError("Must specify unique names.")
        self.names = names
        self.dim = len(names)
        self.coords = tuple(Coordinate(name, cs=self) for name in names)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[self.names.index(key)]
        return self.coords[key]

    def __repr__(self):
        return "{" + ",".join(self.names) + "}"

    def unit_vector_fields(self, dist):
        """Constant unit vector fields e_i (reference coords.py:183-189)."""
        fields = []
        for i, c in enumerate(self.coords):
            ec = dist.VectorField(self, name=f"e{c.name}")
            ec['c'][i] = 1
            fields.append(ec)
        return tuple(fields)


class AffineCOV:
    """Affine change of variables native<->problem interval (reference basis.py:60-93)."""

    def __init__(self, native_bounds, problem_bounds):
        self.native_bounds = native_bounds
        self.problem_bounds = problem_bounds
        self.native_left, self.native_right = native_bounds
        self.native_length = native_bounds[1] - native_bounds[0]
        self.problem_left, self.problem_right = problem_bounds
        self.problem_length = problem_bounds[1] - problem_bounds[0]
        self.native_center = (native_bounds[0] + native_bounds[1]) / 2
        self.problem_center = (problem_bounds[0] + problem_bounds[1]) / 2
        self.stretch = self.problem_length / self.native_length

    def problem_coord(self, native_coord):
        if isinstance(native_coord, str):
            return {'left': self.problem_left, 'right': self.problem_right, 'center': self.problem_center}[native_coord]
        neutral = (np.asarray(native_coord) - self.native_left) / self.native_length
        return self.problem_left + neutral * self.problem_length

    def native_coord(self, problem_coord):
        if isinstance(problem_coord, str):
            return {'left': self.native_left, 'right': self.native_right, 'center': self.native_center}[problem_coord]
        neutral = (np.asarray(problem_coord) - self.problem_left) / self.problem_length
        return self.native_left + neutral * self.native_length


class AzimuthalCoordinate(Coordinate):
    pass


class S2Coordinates:
    """S2 coordinate system (azimuth, colatitude); spin component ordering (-, +) (reference coords.py:201-252)."""
    spin_ordering = (-1, +1)
    dim = 2
    curvilinear = True

    def __init__(self, azimuth, colatitude):
        self.names = (azimuth, colatitude)
        self.azimuth = AzimuthalCoordinate(azimuth, cs=self)
        self.colatitude = Coordinate(colatitude, cs=self)
        self.coords = (self.azimuth, self.colatitude)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[self.names.index(key)]
        return self.coords[key]

    def __repr__(self):
        return "{" + ",".join(self.names) + "}"

    @classmethod
    def U_forward(cls, order=1):
        """Unitary map from coordinate (phi, theta) to spin (-, +) components: u[+-] = (u[theta] +- 1j u[phi]) / sqrt(2)
        (reference coords.py:219-227)."""
        Ui = {+1: np.array([+1j, 1]) / np.sqrt(2), -1: np.array([-1j, 1]) / np.sqrt(2)}
        U = np.array([Ui[s] for s in cls.spin_ordering])
        out = U
        for _ in range(order - 1):
            out = np.kron(out, U)
        return out

    @classmethod
    def U_backward(cls, order=1):
        return cls.U_forward(order).T.conj()


class SphericalCoordinates:
    """Spherical coordinate system (azimuth, colatitude, radius); spin and regularity component ordering (-, +, 0)
    (reference coords.py:313-385)."""
    spin_ordering = (-1, +1, 0)
    reg_ordering = (-1, +1, 0)
    dim = 3
    curvilinear = True

    def __init__(self, azimuth, colatitude, radius):
        self.names = (azimuth, colatitude, radius)
        self.azimuth = AzimuthalCoordinate(azimuth, cs=self)
        self.colatitude = Coordinate(colatitude, cs=self)
        self.radius = Coordinate(radius, cs=self)
        self.S2coordsys = S2Coordinates(azimuth, colatitude)
        # the sphere's coordinate system shares the coordinate OBJECTS (a sphere basis inside a 3-D distributor sits on them)
        self.S2coordsys.azimuth, self.S2coordsys.colatitude = self.azimuth, self.colatitude
        self.S2coordsys.coords = (self.azimuth, self.colatitude)
        self.coords = (self.azimuth, self.colatitude, self.radius)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[self.names.index(key)]
        return self.coords[key]

    def __repr__(self):
        return "{" + ",".join(self.names) + "}"

    @classmethod
    def U_forward(cls, order=1):
        """Unitary map from coordinate (phi, theta, r) to spin (-, +, 0) components (reference coords.py:336-345)."""
        Ui = {+1: np.array([+1j, 1, 0]) / np.sqrt(2), -1: np.array([-1j, 1, 0]) / np.sqrt(2), 0: np.array([0, 0, 1])}
        U = np.array([Ui[s] for s in cls.spin_ordering])
        out = U
        for _ in range(order - 1):
            out = np.kron(out, U)
        return out

    @classmethod
    def U_backward(cls, order=1):
        return cls.U_forward(order).T.conj()

    @staticmethod
    def cartesian(phi, theta, r):
        return r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(theta)
