"""Unitree Go2 rigid-body model constants and the fp32 model blob the engine consumes.

The numbers restate (not copy) the reference's robot description:
  * spigym/data/robots/go2/urdf/go2.urdf                (masses, inertias, joint frames, foot sphere)
  * spigym/config/robot/go2/go2.yaml:29-128             (dof/body order, limits, default pose, PD gains,
                                                          action_scale / action_clip_value / clip_torques)
  * spigym/config/simulator/isaacgym.yaml:14-28         (200 Hz physics, control_decimation 4)
  * spigym/simulator/isaacgym/isaacgym.py:69-71         (gravity (0,0,-9.81), z up)
  * spigym/config/terrain/plane.yaml:12-14              (plane friction 1.0)

`model_from_urdf` parses a user-supplied go2.urdf with the stdlib XML parser so the embedded table can
be checked against (or replaced by) the file the reference ships; tests do exactly that when the
reference checkout is present.

Blob layout: include/spi_b200.h (SPI_BLOB_*).
"""
from __future__ import annotations

import copy
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np

# ---- blob offsets (mirror of include/spi_b200.h; tests parse the header and compare) -----------------
BLOB = dict(
    MAGIC=0, DT=1, GRAVITY_Z=2, ACTION_SCALE=3, ACTION_CLIP=4, CONTACT_KN=5, CONTACT_CN=6,
    CONTACT_MU=7, CONTACT_DT=8, FOOT_RADIUS=9, NSUB=10, CONTACT_VEPS=11, FOOT_SPHERE=12, BASE_INERTIAL=16,
    BASE_LUMPS=26, LEG_BODIES=46, FOOT_OFFSET=214, Q_DEFAULT=226, TORQUE_LIMIT=238, KP=250, KD=262,
    Q_LOWER=274, Q_UPPER=286, QD_LIMIT=298, SIZE=312,
)
BLOB_MAGIC_VALUE = 20025.0
LEG_BODY_STRIDE = 14
INERTIAL_STRIDE = 10

# candidate parameter ids (mirror of SPI_PARAM_*)
PARAM_IDS: Dict[str, int] = dict(
    mass=0, comx=1, comy=2, comz=3, inertiax=4, inertiay=5, inertiaz=6,
    inertiaxy=7, inertiaxz=8, inertiayz=9,
    motor_model_hip_a=10, motor_model_thigh_a=11, motor_model_calf_a=12, mass_scale=13,
)
# aliases used by act2tau_scalar / act2tau_vec3 kwargs (active_sysid_openloop.py:356-379)
PARAM_IDS.update(scalar_gain=10, hip_gain=10, thigh_gain=11, calf_gain=12)

MOTOR_MODELS = dict(none=0, act2tau_scalar=1, act2tau_vec3=2, act2tau_vec3_tanh=3)
FLAG_HIP_HALF = 1 << 0
FLAG_INERTIA_KEEP = 1 << 1
FLAG_STRICT_INERTIAY = 1 << 2
FLAG_TANH_BEFORE_CLIP = 1 << 3

STATE_DIM = 37
TARGET_DIM = 19
NQ = 12

LEGS = ("FL", "FR", "RL", "RR")
DOF_NAMES = [f"{leg}_{j}_joint" for leg in LEGS for j in ("hip", "thigh", "calf")]
# Isaac Gym body order of go2.yaml:44 (19 bodies; feet and heads kept by dont_collapse)
BODY_NAMES = [
    "base", "FL_hip", "FL_thigh", "FL_calf", "FL_foot", "FR_hip", "FR_thigh", "FR_calf", "FR_foot",
    "Head_upper", "Head_lower", "RL_hip", "RL_thigh", "RL_calf", "RL_foot", "RR_hip", "RR_thigh",
    "RR_calf", "RR_foot",
]


@dataclass
class Inertial:
    mass: float
    com: Sequence[float]            # in link frame
    inertia: Sequence[float]        # (xx, yy, zz, xy, xz, yz) about the com, link axes

    def as_row(self) -> List[float]:
        return [float(self.mass), *map(float, self.com), *map(float, self.inertia)]

    def matrix(self) -> np.ndarray:
        xx, yy, zz, xy, xz, yz = self.inertia
        return np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]], dtype=np.float64)


def _skew(v):
    x, y, z = v
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=np.float64)


def lump(parent: Inertial, child: Inertial, child_pos: Sequence[float]) -> Inertial:
    """Rigidly attach `child` (its link frame at `child_pos` in the parent frame, same orientation)
    to `parent`: parallel-axis combination about the new common centre of mass."""
    mp, mc = float(parent.mass), float(child.mass)
    cp = np.asarray(parent.com, dtype=np.float64)
    cc = np.asarray(child_pos, dtype=np.float64) + np.asarray(child.com, dtype=np.float64)
    m = mp + mc
    if m <= 0.0:
        return copy.deepcopy(parent)
    c = (mp * cp + mc * cc) / m
    I = np.zeros((3, 3))
    for mass, com, Ic in ((mp, cp, parent.matrix()), (mc, cc, child.matrix())):
        d = com - c
        I += Ic + mass * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
    return Inertial(m, c.tolist(), [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]])


@dataclass
class ContactParams:
    """Compliant sphere-plane foot contact of the B200 engine.  These are OUR
    model's constants (PhysX's rigid TGS contact cannot be restated); they are part of the blob
    so a user can retune them without rebuilding."""
    kn: float = 10000.0     # N/m
    cn: float = 0.5         # s/m   Hunt-Crossley: f_n = kn*d*(1 - cn*vz), clamped >= 0
    mu: float = 1.0         # plane.yaml static/dynamic friction
    dt: float = 30.0        # N s/m viscous tangential coefficient, capped by mu*f_n/|v_t|
    veps: float = 1e-3      # m/s
    nsub: int = 2           # integrator sub-steps per 5 ms physics step


@dataclass
class Go2Model:
    base: Inertial
    base_lumps: List[tuple]                 # [(Inertial, pos[3])]  fixed children kept separate (heads)
    leg_bodies: List[Inertial]              # 12, order FL hip,thigh,calf, FR ..., feet NOT yet lumped
    feet: List[Inertial]                    # 4
    joint_origin: List[Sequence[float]]     # 12, in parent frame
    joint_axis: List[int]                   # 12, 0 = x, 1 = y
    foot_joint_origin: List[Sequence[float]]  # 4, foot link frame in calf frame
    foot_sphere_offset: Sequence[float] = (-0.002, 0.0, 0.0)
    foot_radius: float = 0.022
    q_default: Sequence[float] = field(default_factory=list)
    torque_limit: Sequence[float] = field(default_factory=list)
    kp: Sequence[float] = field(default_factory=list)
    kd: Sequence[float] = field(default_factory=list)
    q_lower: Sequence[float] = field(default_factory=list)
    q_upper: Sequence[float] = field(default_factory=list)
    qd_limit: Sequence[float] = field(default_factory=list)
    dt: float = 0.005               # 1 / fps (isaacgym.yaml:15)
    control_decimation: int = 4     # isaacgym.yaml:16
    gravity_z: float = -9.81
    action_scale: float = 0.25      # go2.yaml:107
    action_clip: float = 20.0       # go2.yaml:108
    contact: ContactParams = field(default_factory=ContactParams)

    # ---- derived ------------------------------------------------------------------------------
    def body_masses_isaac_order(self) -> np.ndarray:
        """19 body masses in go2.yaml:44 order (what capture_reference_masses returns,
        scripts/eval.py:185-189)."""
        out = []
        for name in BODY_NAMES:
            if name == "base":
                out.append(self.base.mass)
            elif name.startswith("Head"):
                out.append(self.base_lumps[0 if name == "Head_upper" else 1][0].mass)
            else:
                leg, part = name.split("_")
                li = LEGS.index(leg)
                if part == "foot":
                    out.append(self.feet[li].mass)
                else:
                    out.append(self.leg_bodies[3 * li + ("hip", "thigh", "calf").index(part)].mass)
        return np.asarray(out, dtype=np.float32)

    def total_mass(self) -> float:
        return float(self.body_masses_isaac_order().astype(np.float64).sum())

    def lumped_leg_bodies(self) -> List[Inertial]:
        out = []
        for i, b in enumerate(self.leg_bodies):
            if i % 3 == 2:
                li = i // 3
                b = lump(b, self.feet[li], self.foot_joint_origin[li])
            out.append(b)
        return out

    def lumped_base(self, base: Inertial | None = None) -> Inertial:
        b = copy.deepcopy(self.base if base is None else base)
        for child, pos in self.base_lumps:
            b = lump(b, child, pos)
        return b


def _mirror(i: Inertial, sy: float) -> Inertial:
    """URDF right-side links mirror the left ones across the xz-plane."""
    xx, yy, zz, xy, xz, yz = i.inertia
    return Inertial(i.mass, [i.com[0], sy * i.com[1], i.com[2]], [xx, yy, zz, sy * xy, xz, sy * yz])


def go2_nominal(contact: ContactParams | None = None) -> Go2Model:
    """The Go2 of go2.urdf / go2.yaml, typed in from the URDF numbers (see module docstring)."""
    base = Inertial(6.921, [0.021112, 0.0, -0.005366],
                    [0.02448, 0.098077, 0.107, 0.00012166, 0.0014849, -3.12e-05])
    head_upper = (Inertial(0.001, [0, 0, 0], [9.6e-06, 9.6e-06, 9.6e-06, 0, 0, 0]), [0.285, 0.0, 0.01])
    # Head_lower hangs off Head_upper at (0.008, 0, -0.07)
    head_lower = (Inertial(0.001, [0, 0, 0], [9.6e-06, 9.6e-06, 9.6e-06, 0, 0, 0]), [0.293, 0.0, -0.06])

    hip_fl = Inertial(0.678, [-0.0054, 0.00194, -0.000105],
                      [0.00048, 0.000884, 0.000596, -3.01e-06, 1.11e-06, -1.42e-06])
    thigh_l = Inertial(1.152, [-0.00374, -0.0223, -0.0327],
                       [0.00584, 0.0058, 0.00103, 8.72e-05, -0.000289, 0.000808])
    calf_l = Inertial(0.154, [0.00548, -0.000975, -0.115],
                      [0.00108, 0.0011, 3.29e-05, 3.4e-07, 1.72e-05, 8.28e-06])
    foot = Inertial(0.04, [0, 0, 0], [9.6e-06, 9.6e-06, 9.6e-06, 0, 0, 0])

    leg_bodies, joint_origin, joint_axis, feet, foot_origin = [], [], [], [], []
    for leg in LEGS:
        front = leg[0] == "F"
        left = leg[1] == "L"
        sy = 1.0 if left else -1.0
        sx = 1.0 if front else -1.0
        # hip: rear hips mirror the com x and the xy/xz products (urdf RL_hip / RR_hip)
        xx, yy, zz, xy, xz, yz = hip_fl.inertia
        hip = Inertial(hip_fl.mass, [sx * hip_fl.com[0], sy * hip_fl.com[1], hip_fl.com[2]],
                       [xx, yy, zz, sx * sy * xy, sx * xz, sy * yz])
        leg_bodies += [hip, _mirror(thigh_l, sy), _mirror(calf_l, sy)]
        joint_origin += [[sx * 0.1934, sy * 0.0465, 0.0], [0.0, sy * 0.0955, 0.0], [0.0, 0.0, -0.213]]
        joint_axis += [0, 1, 1]
        feet.append(copy.deepcopy(foot))
        foot_origin.append([0.0, 0.0, -0.213])

    q_default = [0.1, 0.8, -1.5, -0.1, 0.8, -1.5, 0.1, 1.0, -1.5, -0.1, 1.0, -1.5]
    return Go2Model(
        base=base, base_lumps=[head_upper, head_lower], leg_bodies=leg_bodies, feet=feet,
        joint_origin=joint_origin, joint_axis=joint_axis, foot_joint_origin=foot_origin,
        q_default=q_default,
        torque_limit=[23.7, 23.7, 35.55] * 4,
        kp=[25.0] * 12, kd=[0.6] * 12,
        q_lower=[-1.0472, -1.5708, -2.7227, -1.0472, -1.5708, -2.7227,
                 -1.0472, -0.5236, -2.7227, -1.0472, -0.5236, -2.7227],
        q_upper=[1.0472, 3.4907, -0.83776, 1.0472, 3.4907, -0.83776,
                 1.0472, 4.5379, -0.83776, 1.0472, 4.5379, -0.83776],
        qd_limit=[30.1, 30.1, 20.07] * 4,
        contact=contact or ContactParams(),
    )


def model_from_urdf(path, contact: ContactParams | None = None) -> Go2Model:
    """Parse a go2.urdf (the reference ships one at spigym/data/robots/go2/urdf/go2.urdf) into the
    13-body tree.  Fixed joints: feet are kept as separate lumps of the calves, Head_upper/Head_lower
    as lumps of the base; mass-less links (calflower*, imu, radar) carry no inertia and are dropped,
    mirroring `collapse_fixed_joints: True` (go2.yaml:112)."""
    root = ET.parse(str(path)).getroot()

    def inertial_of(name):
        link = next(l for l in root.findall("link") if l.attrib["name"] == name)
        i = link.find("inertial")
        xyz = [float(v) for v in i.find("origin").attrib["xyz"].split()]
        a = i.find("inertia").attrib
        return Inertial(float(i.find("mass").attrib["value"]), xyz,
                        [float(a[k]) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")])

    joints = {j.attrib["name"]: j for j in root.findall("joint")}

    def origin(jname):
        return [float(v) for v in joints[jname].find("origin").attrib["xyz"].split()]

    def axis_id(jname):
        a = [float(v) for v in joints[jname].find("axis").attrib["xyz"].split()]
        assert sorted(map(abs, a)) == [0.0, 0.0, 1.0] and max(a) == 1.0, f"unsupported axis {a}"
        return a.index(1.0)

    m = go2_nominal(contact)
    m.base = inertial_of("base")
    hu, hl = origin("Head_upper_joint"), origin("Head_lower_joint")
    m.base_lumps = [(inertial_of("Head_upper"), hu),
                    (inertial_of("Head_lower"), [hu[i] + hl[i] for i in range(3)])]
    m.leg_bodies, m.joint_origin, m.joint_axis, m.feet, m.foot_joint_origin = [], [], [], [], []
    lo, hi, eff, vel = [], [], [], []
    for leg in LEGS:
        for part in ("hip", "thigh", "calf"):
            jn = f"{leg}_{part}_joint"
            m.leg_bodies.append(inertial_of(f"{leg}_{part}"))
            m.joint_origin.append(origin(jn))
            m.joint_axis.append(axis_id(jn))
            lim = joints[jn].find("limit").attrib
            lo.append(float(lim["lower"])); hi.append(float(lim["upper"]))
            eff.append(float(lim["effort"])); vel.append(float(lim["velocity"]))
        m.feet.append(inertial_of(f"{leg}_foot"))
        m.foot_joint_origin.append(origin(f"{leg}_foot_joint"))
    m.q_lower, m.q_upper, m.torque_limit, m.qd_limit = lo, hi, eff, vel
    foot_link = next(l for l in root.findall("link") if l.attrib["name"] == "FL_foot")
    col = foot_link.find("collision")
    m.foot_sphere_offset = [float(v) for v in col.find("origin").attrib["xyz"].split()]
    m.foot_radius = float(col.find("geometry").find("sphere").attrib["radius"])
    return m


def build_model_blob(model: Go2Model | None = None) -> np.ndarray:
    """fp32[SPI_BLOB_SIZE] in the layout of include/spi_b200.h."""
    m = model or go2_nominal()
    b = np.zeros(BLOB["SIZE"], dtype=np.float64)
    b[BLOB["MAGIC"]] = BLOB_MAGIC_VALUE
    b[BLOB["DT"]] = m.dt
    b[BLOB["GRAVITY_Z"]] = m.gravity_z
    b[BLOB["ACTION_SCALE"]] = m.action_scale
    b[BLOB["ACTION_CLIP"]] = m.action_clip
    b[BLOB["CONTACT_KN"]] = m.contact.kn
    b[BLOB["CONTACT_CN"]] = m.contact.cn
    b[BLOB["CONTACT_MU"]] = m.contact.mu
    b[BLOB["CONTACT_DT"]] = m.contact.dt
    b[BLOB["FOOT_RADIUS"]] = m.foot_radius
    b[BLOB["NSUB"]] = float(int(m.contact.nsub))
    b[BLOB["CONTACT_VEPS"]] = m.contact.veps
    b[BLOB["FOOT_SPHERE"]:BLOB["FOOT_SPHERE"] + 3] = m.foot_sphere_offset     # foot link frame = FOOT_OFFSET - this
    b[BLOB["BASE_INERTIAL"]:BLOB["BASE_INERTIAL"] + 10] = m.base.as_row()
    for k, (child, pos) in enumerate(m.base_lumps):
        o = BLOB["BASE_LUMPS"] + 10 * k
        # lump record: mass, position of the child's COM in the base frame, inertia about that COM
        cpos = [pos[i] + child.com[i] for i in range(3)]
        b[o:o + 10] = [child.mass, *cpos, *child.inertia]
    for i, body in enumerate(m.lumped_leg_bodies()):
        o = BLOB["LEG_BODIES"] + LEG_BODY_STRIDE * i
        b[o:o + 10] = body.as_row()
        b[o + 10:o + 13] = m.joint_origin[i]
        b[o + 13] = float(m.joint_axis[i])
    for li in range(4):
        o = BLOB["FOOT_OFFSET"] + 3 * li
        b[o:o + 3] = [m.foot_joint_origin[li][k] + m.foot_sphere_offset[k] for k in range(3)]
    for key, arr in (("Q_DEFAULT", m.q_default), ("TORQUE_LIMIT", m.torque_limit), ("KP", m.kp),
                     ("KD", m.kd), ("Q_LOWER", m.q_lower), ("Q_UPPER", m.q_upper),
                     ("QD_LIMIT", m.qd_limit)):
        b[BLOB[key]:BLOB[key] + 12] = arr
    return b.astype(np.float32)


def default_param_vector(model: Go2Model | None = None) -> np.ndarray:
    """Nominal value of every SPI_PARAM_* (URDF base link; motor a = 20.0 as in
    active_sysid_openloop.yaml:25-27; mass_scale 1)."""
    m = model or go2_nominal()
    v = np.zeros(14, dtype=np.float32)
    v[0] = m.base.mass
    v[1:4] = m.base.com
    v[4:10] = m.base.inertia
    v[10:13] = 20.0
    v[13] = 1.0
    return v
