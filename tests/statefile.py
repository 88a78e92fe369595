"""Binary state-file format shared by oracle/ref_harness.cpp (reference side) and the tests.

All little-endian.  Node arrays are flat in U order (u = k*ni*nj + j*ni + i); ef is [3*u+c].

  char[8]  "ESPICST1"
  int32    ni, nj, nk, flags (1 = addSphere, 2 = addInlet, 4 = keep phi from geometry setup), nsp, pad
  double   x0[3], xm[3], dt
  double   sphere_c[3], sphere_r, sphere_phi
  double   phi0, Te0, n0
  double   phi[nn], rho[nn], ef[3nn], node_vol[nn];  int32 object_id[nn]
  per species: double mass, charge, mpw0; int64 np; double den[nn], den_ave[nn]; double part[7][np]
  double   diag[16]: converged, PE, then per species (first two): real_count, px, py, pz, KE
"""
import os
import struct
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class State:
    def __init__(self):
        self.ni = self.nj = self.nk = 0
        self.flags = 0
        self.x0 = np.zeros(3)
        self.xm = np.zeros(3)
        self.dt = 0.0
        self.sphere_c = np.zeros(3)
        self.sphere_r = 0.0
        self.sphere_phi = 0.0
        self.phi0, self.Te0, self.n0 = 0.0, 1.5, 1e12
        self.phi = self.rho = self.ef = self.node_vol = self.object_id = None
        self.species = []   # dicts: mass, charge, mpw0, den, den_ave, part (7,np)
        self.diag = np.zeros(16)

    @property
    def nn(self):
        return self.ni * self.nj * self.nk


def write_state(path, s):
    nn = s.nn
    with open(path, "wb") as f:
        f.write(b"ESPICST1")
        f.write(struct.pack("<6i", s.ni, s.nj, s.nk, s.flags, len(s.species), 0))
        f.write(np.asarray(s.x0, dtype="<f8").tobytes())
        f.write(np.asarray(s.xm, dtype="<f8").tobytes())
        f.write(struct.pack("<d", s.dt))
        f.write(np.asarray(s.sphere_c, dtype="<f8").tobytes())
        f.write(struct.pack("<2d", s.sphere_r, s.sphere_phi))
        f.write(struct.pack("<3d", s.phi0, s.Te0, s.n0))
        for a, n in ((s.phi, nn), (s.rho, nn), (s.ef, 3 * nn), (s.node_vol, nn)):
            a = np.zeros(n) if a is None else np.asarray(a, dtype="<f8")
            assert a.size == n
            f.write(a.tobytes())
        oid = np.zeros(nn, dtype="<i4") if s.object_id is None else np.asarray(s.object_id, dtype="<i4")
        f.write(oid.tobytes())
        for sp in s.species:
            part = np.ascontiguousarray(sp["part"], dtype="<f8")
            f.write(struct.pack("<3dq", sp["mass"], sp["charge"], sp.get("mpw0", 1.0), part.shape[1]))
            for key in ("den", "den_ave"):
                a = sp.get(key)
                a = np.zeros(nn) if a is None else np.asarray(a, dtype="<f8")
                f.write(a.tobytes())
            f.write(part.tobytes())
        f.write(np.asarray(s.diag, dtype="<f8").tobytes())


def read_state(path):
    s = State()
    with open(path, "rb") as f:
        assert f.read(8) == b"ESPICST1"
        s.ni, s.nj, s.nk, s.flags, nsp, _ = struct.unpack("<6i", f.read(24))
        rd = lambda n: np.frombuffer(f.read(8 * n), dtype="<f8").copy()
        s.x0, s.xm = rd(3), rd(3)
        s.dt = rd(1)[0]
        s.sphere_c = rd(3)
        s.sphere_r, s.sphere_phi = rd(2)
        s.phi0, s.Te0, s.n0 = rd(3)
        nn = s.nn
        s.phi, s.rho, s.ef, s.node_vol = rd(nn), rd(nn), rd(3 * nn), rd(nn)
        s.object_id = np.frombuffer(f.read(4 * nn), dtype="<i4").copy()
        for _ in range(nsp):
            mass, charge, mpw0, np_ = struct.unpack("<3dq", f.read(32))
            den, den_ave = rd(nn), rd(nn)
            part = rd(7 * np_).reshape(7, np_)
            s.species.append(dict(mass=mass, charge=charge, mpw0=mpw0, den=den, den_ave=den_ave, part=part))
        s.diag = rd(16)
    return s


def have_ref(which="ref_ch3"):
    return os.path.exists(os.path.join(REF_DIR, which))


def run_ref(which, state, cmds, tmpdir):
    """Run an oracle/_ref harness binary (built from the unmodified reference) on `state`."""
    fin = os.path.join(str(tmpdir), "in.state")
    fout = os.path.join(str(tmpdir), "out.state")
    write_state(fin, state)
    out = subprocess.run([os.path.join(REF_DIR, which), fin, fout] + list(cmds), check=True,
                         capture_output=True, text=True)
    res = read_state(fout)
    res.stdout = out.stdout
    res.stderr = out.stderr
    return res


def state_from_oracle(world, species, dt, flags=None):
    """Snapshot an oracle.World (+ oracle.Species list) into a State."""
    s = State()
    s.ni, s.nj, s.nk = world.ni, world.nj, world.nk
    s.x0, s.xm, s.dt = world.x0, world.xm, dt
    f = 0
    if world.sphere is not None:
        f |= 1
        s.sphere_c = np.array(world.sphere[0])
        s.sphere_r, s.sphere_phi = world.sphere[1], world.sphere[2]
    if world.inlet:
        f |= 2
    s.flags = f if flags is None else flags
    s.phi0, s.Te0, s.n0 = world.phi0, world.Te0, world.n0
    s.phi, s.rho, s.ef = world.phi.copy(), world.rho.copy(), world.ef.copy()
    s.node_vol, s.object_id = world.node_vol.copy(), world.object_id.copy()
    for sp in species:
        s.species.append(dict(mass=sp.mass, charge=sp.charge, mpw0=sp.mpw0, den=sp.den.copy(),
                              den_ave=sp.den_ave.copy(), part=sp.particles()))
    return s


# ---- npz (golden fixture) round trip -------------------------------------------------------

_SCALARS = ("ni", "nj", "nk", "flags", "dt", "sphere_r", "sphere_phi", "phi0", "Te0", "n0")
_ARRAYS = ("x0", "xm", "sphere_c", "phi", "rho", "ef", "node_vol", "object_id", "diag")


def state_to_dict(st, prefix):
    d = {}
    for k in _SCALARS:
        d[prefix + k] = np.array(getattr(st, k))
    for k in _ARRAYS:
        d[prefix + k] = np.asarray(getattr(st, k))
    d[prefix + "nsp"] = np.array(len(st.species))
    for q, sp in enumerate(st.species):
        d["%ssp%d_params" % (prefix, q)] = np.array([sp["mass"], sp["charge"], sp.get("mpw0", 1.0)])
        for k in ("den", "den_ave", "part"):
            d["%ssp%d_%s" % (prefix, q, k)] = np.asarray(sp[k])
    return d


def state_from_dict(d, prefix):
    st = State()
    for k in _SCALARS:
        v = d[prefix + k][()]
        setattr(st, k, int(v) if k in ("ni", "nj", "nk", "flags") else float(v))
    for k in _ARRAYS:
        setattr(st, k, np.array(d[prefix + k]))
    for q in range(int(d[prefix + "nsp"][()])):
        m, c, w0 = d["%ssp%d_params" % (prefix, q)]
        st.species.append(dict(mass=float(m), charge=float(c), mpw0=float(w0),
                               den=np.array(d["%ssp%d_den" % (prefix, q)]),
                               den_ave=np.array(d["%ssp%d_den_ave" % (prefix, q)]),
                               part=np.array(d["%ssp%d_part" % (prefix, q)])))
    return st
