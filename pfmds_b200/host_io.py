"""Python readers for the reference's file formats (tests, tools); the product host reads them in C++ (host/md_inputs.hpp)."""
import numpy as np


def read_xyz(path):
    """extended-xyz as read_box_size/read_particles do (md_read_write.f90:22-61)."""
    with open(path) as f:
        n = int(f.readline().split()[0])
        t = f.readline().replace(",", " ").split()
        m = [float(x) for x in t[1:10]]
        pos, vel, mass, names = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), []
        for i in range(n):
            a = f.readline().split()
            pos[i] = [float(x) for x in a[0:3]]
            vel[i] = [float(x) for x in a[3:6]]
            mass[i] = float(a[6])
            names.append(a[7])
    return dict(box=np.array([m[0], m[4], m[8]]), pos=pos, vel=vel, mass=mass, names=names)
