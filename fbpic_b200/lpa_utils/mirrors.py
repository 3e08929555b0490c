"""`Mirror`: zero E and B in a thin slab orthogonal to z at every step (fbpic/lpa_utils/mirrors.py:10-94).
On the device the slab is a contiguous block of rows of each [Nz][Nr] array: one memset per array."""
from scipy.constants import c


class Mirror(object):

    def __init__(self, z_start, z_end, gamma_boost=None, m='all'):
        self.z_start, self.z_end, self.gamma_boost = z_start, z_end, gamma_boost
        if m == 'all':
            self.modes = None
        elif isinstance(m, int):
            self.modes = [m]
        elif isinstance(m, list):
            self.modes = m
        else:
            raise TypeError('m should be an int or a list of ints.')

    def set_fields_to_zero(self, interp, comm, t_boost):
        """mirrors.py:46-94"""
        if self.gamma_boost is None:
            z0, z1 = self.z_start, self.z_end
        else:
            beta = (1. - 1. / self.gamma_boost**2)**.5
            z0 = 1. / self.gamma_boost * self.z_start - beta * c * t_boost
            z1 = 1. / self.gamma_boost * self.z_end - beta * c * t_boost
        zmin, zmax = comm.get_zmin_zmax(local=True, with_guard=True, with_damp=True, rank=comm.rank)
        if (z0 < zmin) or (z0 >= zmax):
            return
        imax = int((z0 - zmin) / interp[0].dz)
        n_cells = int((z1 - z0) / interp[0].dz)
        imin = max(imax - n_cells, 0)
        if imax <= imin:
            return
        for i, grid in enumerate(interp):
            if self.modes is not None and i not in self.modes:
                continue
            names = ['Er', 'Et', 'Ez', 'Br', 'Bt', 'Bz']
            if grid.use_pml:
                names += ['Er_pml', 'Et_pml', 'Br_pml', 'Bt_pml']
            for k in names:
                arr = getattr(grid, k)
                if hasattr(arr, 'ptr'):
                    arr[imin:imax].fill(0)        # DeviceArray row range: cudaMemsetAsync
                else:
                    arr[imin:imax, :] = 0.
