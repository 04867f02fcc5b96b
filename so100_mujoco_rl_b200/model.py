"""Host-side MJCF reader for the so100 scene -> raw model constants (fp64).

Replaces, for the hot path only, the model load of the reference
(`mujoco.MjModel.from_xml_path` at src/so100_mujoco_rl/envs/env_base_01.py:38 and `joints_from_model`,
src/so100_mujoco_rl/envs/utils.py:64-89).  It understands exactly the MJCF subset the so100 scene uses:
`<compiler angle>`, nested `<default class>` with `childclass`, `<body pos quat|euler>`, `<inertial>`,
hinge `<joint class>`, `<position>` actuators, one `<camera>`, and the scene's `<attach model body prefix>`.
Nothing is derived here (no dof_M0 / kv): the C library and the oracle each derive those themselves.

Two inputs are accepted:
  * the committed mesh-free flat scene  `assets/so100_scene.xml`  (default; always available), and
  * the reference's own two-file form  `env01.xml` (+ `so_arm100_camera.xml` through `<attach>`), when present.
"""
from __future__ import annotations

import ctypes
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

NJ = 6
MAX_PAD = 8
PREFIX = "so100_"  # reference: MUJOCO_SO100_PREFIX, envs/utils.py:7
JOINT_NAMES = ["Rotation", "Pitch", "Elbow", "Wrist_Pitch", "Wrist_Roll", "Jaw"]
BODY_NAMES = ["Rotation_Pitch", "Upper_Arm", "Lower_Arm", "Wrist_Pitch_Roll", "Fixed_Jaw", "Moving_Jaw"]
CAMERA_NAME = "so100_end_point_camera"  # envs/utils.py:99
ASSET_SCENE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "so100_scene.xml")

# MuJoCo defaults for elements the scene does not override
_DEF_SOLREF = (0.02, 1.0)
_DEF_SOLIMP = (0.9, 0.95, 0.001, 0.5, 2.0)


def _floats(s: str) -> list[float]:
    return [float(x) for x in s.split()]


def quat_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ])


def euler_to_quat(e, seq="xyz"):
    """MuJoCo eulerseq semantics: lower-case = intrinsic (q <- q * q_axis), upper-case = extrinsic."""
    q = np.array([1.0, 0.0, 0.0, 0.0])
    for ang, ch in zip(e, seq):
        qa = np.zeros(4)
        qa[0] = math.cos(ang / 2)
        qa["xyz".index(ch.lower()) + 1] = math.sin(ang / 2)
        q = quat_mul(q, qa) if ch.islower() else quat_mul(qa, q)
    return q


def quat_to_mat(q):
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


class So100Model(ctypes.Structure):
    """ctypes mirror of `so100_model` in include/so100_b200.h (field order and types must match)."""
    _fields_ = [
        ("struct_size", ctypes.c_int32),
        ("nsubstep", ctypes.c_int32),
        ("timestep", ctypes.c_double),
        ("gravity", ctypes.c_double * 3),
        ("base_pos", ctypes.c_double * 3),
        ("base_quat", ctypes.c_double * 4),
        ("body_pos", (ctypes.c_double * 3) * NJ),
        ("body_quat", (ctypes.c_double * 4) * NJ),
        ("body_ipos", (ctypes.c_double * 3) * NJ),
        ("body_iquat", (ctypes.c_double * 4) * NJ),
        ("body_mass", ctypes.c_double * NJ),
        ("body_inertia", (ctypes.c_double * 3) * NJ),
        ("jnt_axis", (ctypes.c_double * 3) * NJ),
        ("jnt_range", (ctypes.c_double * 2) * NJ),
        ("jnt_armature", ctypes.c_double * NJ),
        ("jnt_frictionloss", ctypes.c_double * NJ),
        ("jnt_solref_limit", (ctypes.c_double * 2) * NJ),
        ("jnt_solimp_limit", (ctypes.c_double * 5) * NJ),
        ("dof_solref_friction", (ctypes.c_double * 2) * NJ),
        ("dof_solimp_friction", (ctypes.c_double * 5) * NJ),
        ("act_kp", ctypes.c_double * NJ),
        ("act_dampratio", ctypes.c_double * NJ),
        ("act_kv", ctypes.c_double * NJ),
        ("act_ctrlrange", (ctypes.c_double * 2) * NJ),
        ("act_forcerange", (ctypes.c_double * 2) * NJ),
        ("ee_body", ctypes.c_int32),
        ("wrist_body", ctypes.c_int32),
        ("cam_body", ctypes.c_int32),
        ("_pad0", ctypes.c_int32),
        ("ee_offset", ctypes.c_double * 3),
        ("cam_pos", ctypes.c_double * 3),
        ("cam_quat", ctypes.c_double * 4),
        ("cam_fovy_deg", ctypes.c_double),
        ("block_half_z", ctypes.c_double),
        ("block_mass", ctypes.c_double),
        ("block_friction", ctypes.c_double),
        ("contact_solref", ctypes.c_double * 2),
        ("contact_solimp", ctypes.c_double * 5),
        ("block_ncon", ctypes.c_int32),
        ("_pad1", ctypes.c_int32),
        ("n_pad", ctypes.c_int32),
        ("_pad2", ctypes.c_int32),
        ("pad_body", ctypes.c_int32 * MAX_PAD),
        ("pad_pos", (ctypes.c_double * 3) * MAX_PAD),
        ("pad_size", (ctypes.c_double * 3) * MAX_PAD),
        ("pad_solref", ctypes.c_double * 2),
        ("pad_solimp", ctypes.c_double * 5),
        ("pad_friction", ctypes.c_double),
        ("floor_solref", ctypes.c_double * 2),
        ("floor_solimp", ctypes.c_double * 5),
        ("floor_friction", ctypes.c_double),
    ]


@dataclass
class ModelSpec:
    """Raw so100 constants; arrays are indexed by joint/body 0..5 (Rotation .. Jaw)."""
    timestep: float = 0.002
    nsubstep: int = 16  # frame_skip, env_base_01.py:45
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    base_pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
    base_quat: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
    body_pos: np.ndarray = field(default_factory=lambda: np.zeros((NJ, 3)))
    body_quat: np.ndarray = field(default_factory=lambda: np.tile([1.0, 0, 0, 0], (NJ, 1)))
    body_ipos: np.ndarray = field(default_factory=lambda: np.zeros((NJ, 3)))
    body_iquat: np.ndarray = field(default_factory=lambda: np.tile([1.0, 0, 0, 0], (NJ, 1)))
    body_mass: np.ndarray = field(default_factory=lambda: np.zeros(NJ))
    body_inertia: np.ndarray = field(default_factory=lambda: np.zeros((NJ, 3)))
    jnt_axis: np.ndarray = field(default_factory=lambda: np.zeros((NJ, 3)))
    jnt_range: np.ndarray = field(default_factory=lambda: np.zeros((NJ, 2)))
    jnt_armature: np.ndarray = field(default_factory=lambda: np.zeros(NJ))
    jnt_frictionloss: np.ndarray = field(default_factory=lambda: np.zeros(NJ))
    jnt_solref_limit: np.ndarray = field(default_factory=lambda: np.tile(_DEF_SOLREF, (NJ, 1)))
    jnt_solimp_limit: np.ndarray = field(default_factory=lambda: np.tile(_DEF_SOLIMP, (NJ, 1)))
    dof_solref_friction: np.ndarray = field(default_factory=lambda: np.tile(_DEF_SOLREF, (NJ, 1)))
    dof_solimp_friction: np.ndarray = field(default_factory=lambda: np.tile(_DEF_SOLIMP, (NJ, 1)))
    act_kp: np.ndarray = field(default_factory=lambda: np.zeros(NJ))
    act_dampratio: np.ndarray = field(default_factory=lambda: np.zeros(NJ))
    act_kv: np.ndarray = field(default_factory=lambda: np.zeros(NJ))
    act_ctrlrange: np.ndarray = field(default_factory=lambda: np.zeros((NJ, 2)))
    act_forcerange: np.ndarray = field(default_factory=lambda: np.zeros((NJ, 2)))
    ee_body: int = 4  # Fixed_Jaw
    wrist_body: int = 3  # Wrist_Pitch_Roll
    cam_body: int = 4
    # env_base_01.py:125 builds the offset as a float32 array, so the reference's -0.1 is float32(-0.1) = -0.10000000149
    ee_offset: np.ndarray = field(default_factory=lambda: np.array([0.0, -0.1, 0.0], dtype=np.float32).astype(np.float64))
    cam_pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
    cam_quat: np.ndarray = field(default_factory=lambda: np.array([1.0, 0, 0, 0]))
    cam_fovy_deg: float = 45.0
    # block <-> floor contact (env01.xml:29-34, :39): MuJoCo defaults unless the scene overrides them
    block_half_z: float = 0.01
    block_mass: float = 0.008
    block_friction: float = 1.0
    contact_solref: np.ndarray = field(default_factory=lambda: np.array(_DEF_SOLREF))
    contact_solimp: np.ndarray = field(default_factory=lambda: np.array(_DEF_SOLIMP))
    block_ncon: int = 4  # bottom corners of a flat box on a plane (mjc_PlaneBox); 0 = the pair does not collide
    # arm <-> floor: the jaws' primitive box colliders (so_arm100_camera.xml:60-61, 108-111, 120-123) and the floor geom
    pad_body: list = field(default_factory=list)          # body index per pad (4 = Fixed_Jaw, 5 = Moving_Jaw)
    pad_names: list = field(default_factory=list)
    pad_pos: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    pad_size: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    pad_solref: np.ndarray = field(default_factory=lambda: np.array(_DEF_SOLREF))
    pad_solimp: np.ndarray = field(default_factory=lambda: np.array(_DEF_SOLIMP))
    pad_friction: float = 1.0
    floor_solref: np.ndarray = field(default_factory=lambda: np.array(_DEF_SOLREF))
    floor_solimp: np.ndarray = field(default_factory=lambda: np.array(_DEF_SOLIMP))
    floor_friction: float = 1.0
    joint_names: list = field(default_factory=lambda: list(JOINT_NAMES))
    source: str = ""

    def to_ctypes(self) -> So100Model:
        m = So100Model()
        m.struct_size = ctypes.sizeof(So100Model)
        m.nsubstep = int(self.nsubstep)
        m.timestep = float(self.timestep)
        m.ee_body, m.wrist_body, m.cam_body = int(self.ee_body), int(self.wrist_body), int(self.cam_body)
        m.cam_fovy_deg = float(self.cam_fovy_deg)
        m.block_half_z, m.block_mass, m.block_friction = float(self.block_half_z), float(self.block_mass), float(self.block_friction)
        m.block_ncon = int(self.block_ncon)

        def put(dst, src):
            src = np.asarray(src, dtype=np.float64)
            if src.ndim == 1:
                for i, v in enumerate(src):
                    dst[i] = float(v)
            else:
                for i in range(src.shape[0]):
                    for j in range(src.shape[1]):
                        dst[i][j] = float(src[i, j])

        for name in ("gravity", "base_pos", "base_quat", "body_pos", "body_quat", "body_ipos", "body_iquat",
                     "body_mass", "body_inertia", "jnt_axis", "jnt_range", "jnt_armature", "jnt_frictionloss",
                     "jnt_solref_limit", "jnt_solimp_limit", "dof_solref_friction", "dof_solimp_friction",
                     "act_kp", "act_dampratio", "act_kv", "act_ctrlrange", "act_forcerange", "ee_offset",
                     "cam_pos", "cam_quat", "contact_solref", "contact_solimp"):
            put(getattr(m, name), getattr(self, name))
        m.n_pad = len(self.pad_body)
        for k, b in enumerate(self.pad_body):
            m.pad_body[k] = int(b)
        put(m.pad_pos, np.asarray(self.pad_pos, dtype=np.float64).reshape(-1, 3))
        put(m.pad_size, np.asarray(self.pad_size, dtype=np.float64).reshape(-1, 3))
        for name in ("pad_solref", "pad_solimp", "floor_solref", "floor_solimp"):
            put(getattr(m, name), getattr(self, name))
        m.pad_friction, m.floor_friction = float(self.pad_friction), float(self.floor_friction)
        return m


class _Defaults:
    """Nested <default class=...> resolution: attribute lookup walks from the class up to the root default."""

    def __init__(self):
        self.parent: dict[str, str | None] = {"__root__": None}
        self.attrs: dict[str, dict[str, dict[str, str]]] = {"__root__": {}}

    def load(self, node, parent="__root__", top=True):
        name = node.get("class") if not top or node.get("class") else None
        if top and name is None:
            cls = "__root__"
        else:
            cls = name
            self.parent[cls] = parent
            self.attrs.setdefault(cls, {})
        for ch in node:
            if ch.tag == "default":
                self.load(ch, cls, top=False)
            else:
                self.attrs[cls].setdefault(ch.tag, {}).update(ch.attrib)

    def resolve(self, tag, cls, own: dict[str, str]) -> dict[str, str]:
        chain = []
        c = cls if cls in self.parent else "__root__"
        while c is not None:
            chain.append(c)
            c = self.parent.get(c)
        out: dict[str, str] = {}
        for c in reversed(chain):
            out.update(self.attrs.get(c, {}).get(tag, {}))
        out.update(own)
        return out


def _orientation(el, radians: bool) -> np.ndarray:
    if el.get("quat") is not None:
        q = np.array(_floats(el.get("quat")))
        return q / np.linalg.norm(q)
    if el.get("euler") is not None:
        e = np.array(_floats(el.get("euler")))
        if not radians:
            e = np.deg2rad(e)
        return euler_to_quat(e)
    return np.array([1.0, 0.0, 0.0, 0.0])


def _find_body(root_el, name):
    for b in root_el.iter("body"):
        if b.get("name") == name:
            return b
    return None


def load_model(path: str | None = None) -> ModelSpec:
    """Read the so100 scene.  `path` may be the flat asset (default) or the reference's env01.xml."""
    path = path or ASSET_SCENE
    tree = ET.parse(path)
    root = tree.getroot()
    prefix = ""
    arm_root = root

    attach = None
    for wb in root.findall("worldbody"):
        for a in wb.findall("attach"):
            attach = a
    if attach is not None:
        # reference form: env01.xml:24-26 attaches asset model "so_arm100" (file so_arm100_camera.xml) with a prefix
        model_name = attach.get("model")
        child_file = None
        for asset in root.findall("asset"):
            for mdl in asset.findall("model"):
                if mdl.get("name") == model_name:
                    child_file = mdl.get("file")
        if child_file is None:
            raise ValueError(f"attach refers to unknown model {model_name!r}")
        arm_root = ET.parse(os.path.join(os.path.dirname(path), child_file)).getroot()
        prefix = attach.get("prefix", "")
        base_name = attach.get("body")
    else:
        prefix = PREFIX
        base_name = None

    def rd_compiler(r):
        c = r.find("compiler")
        return (c is not None and c.get("angle", "degree") == "radian")

    radians = rd_compiler(arm_root)

    spec = ModelSpec(source=os.path.abspath(path))
    opt = root.find("option")
    if opt is not None:
        if opt.get("timestep"):
            spec.timestep = float(opt.get("timestep"))
        if opt.get("gravity"):
            spec.gravity = np.array(_floats(opt.get("gravity")))

    defaults = _Defaults()
    for d in arm_root.findall("default"):
        defaults.load(d)

    # locate the Base body of the arm
    def strip(n):
        return n[len(prefix):] if (attach is None and n and n.startswith(prefix)) else n

    base_el = None
    for wb in arm_root.findall("worldbody"):
        for b in wb.iter("body"):
            if strip(b.get("name")) == (base_name or "Base"):
                base_el = b
                break
    if base_el is None:
        raise ValueError("so100 Base body not found")
    if base_el.get("pos"):
        spec.base_pos = np.array(_floats(base_el.get("pos")))
    spec.base_quat = _orientation(base_el, radians)

    childclass = base_el.get("childclass")
    cur = base_el
    joint_names = []
    pad_geoms = []   # (body index, resolved geom attributes) of the arm's primitive (box) colliders
    for i, bname in enumerate(BODY_NAMES):
        nxt = None
        for b in cur.findall("body"):
            if strip(b.get("name")) == bname:
                nxt = b
        if nxt is None:
            raise ValueError(f"body {bname} not found under {cur.get('name')}")
        cur = nxt
        if cur.get("childclass"):
            childclass = cur.get("childclass")
        spec.body_pos[i] = _floats(cur.get("pos", "0 0 0"))
        spec.body_quat[i] = _orientation(cur, radians)
        inert = cur.find("inertial")
        if inert is None:
            raise ValueError(f"body {bname} has no <inertial> (mesh-derived inertia is not supported)")
        spec.body_ipos[i] = _floats(inert.get("pos", "0 0 0"))
        spec.body_iquat[i] = _orientation(inert, radians)
        spec.body_mass[i] = float(inert.get("mass"))
        if inert.get("diaginertia") is None:
            raise ValueError("only diaginertia inertials are supported")
        spec.body_inertia[i] = _floats(inert.get("diaginertia"))
        joints = cur.findall("joint")
        if len(joints) != 1:
            raise ValueError(f"body {bname}: expected exactly one joint")
        j = joints[0]
        ja = defaults.resolve("joint", j.get("class") or childclass, dict(j.attrib))
        if ja.get("type", "hinge") != "hinge":
            raise ValueError("only hinge joints are supported in the arm")
        if any(abs(v) > 0 for v in _floats(ja.get("pos", "0 0 0"))):
            raise ValueError("joint anchors away from the body origin are not supported")
        if float(ja.get("ref", "0")) != 0.0 or float(ja.get("damping", "0")) != 0.0 or float(ja.get("stiffness", "0")) != 0.0:
            raise ValueError("joint ref/damping/stiffness are not supported (so100 uses none)")
        ax = np.array(_floats(ja.get("axis", "0 0 1")))
        spec.jnt_axis[i] = ax / np.linalg.norm(ax)
        if ja.get("range") is None:
            raise ValueError("unlimited joints are not supported")
        rng = np.array(_floats(ja["range"]))
        spec.jnt_range[i] = rng if radians else np.deg2rad(rng)
        spec.jnt_armature[i] = float(ja.get("armature", "0"))
        spec.jnt_frictionloss[i] = float(ja.get("frictionloss", "0"))
        if ja.get("solreflimit"):
            spec.jnt_solref_limit[i] = _floats(ja["solreflimit"])
        if ja.get("solimplimit"):
            v = _floats(ja["solimplimit"])
            spec.jnt_solimp_limit[i, :len(v)] = v
        if ja.get("solreffriction"):
            spec.dof_solref_friction[i] = _floats(ja["solreffriction"])
        if ja.get("solimpfriction"):
            v = _floats(ja["solimpfriction"])
            spec.dof_solimp_friction[i, :len(v)] = v
        joint_names.append(strip(j.get("name")))
        for g in cur.findall("geom"):
            ga = defaults.resolve("geom", g.get("class") or childclass, dict(g.attrib))
            if ga.get("type", "sphere") == "box" and (int(ga.get("contype", "1")) or int(ga.get("conaffinity", "1"))):
                pad_geoms.append((i, ga))
        cam = cur.find("camera")
        if cam is not None and strip(cam.get("name")) in ("end_point_camera",):
            spec.cam_body = i
            spec.cam_pos = np.array(_floats(cam.get("pos", "0 0 0")))
            spec.cam_quat = _orientation(cam, radians)
            spec.cam_fovy_deg = float(cam.get("fovy", "45"))
    spec.joint_names = joint_names
    _read_pads(pad_geoms, defaults, spec)
    spec.ee_body = BODY_NAMES.index("Fixed_Jaw")
    spec.wrist_body = BODY_NAMES.index("Wrist_Pitch_Roll")

    # actuators: <position class= joint=>
    act_el = arm_root.find("actuator")
    seen = set()
    if act_el is not None:
        for a in act_el.findall("position"):
            aa = defaults.resolve("position", a.get("class") or childclass, dict(a.attrib))
            jn = strip(aa.get("joint"))
            if jn not in joint_names:
                continue
            k = joint_names.index(jn)
            seen.add(k)
            spec.act_kp[k] = float(aa.get("kp", "1"))
            if aa.get("kv") is not None:
                spec.act_kv[k] = float(aa["kv"])
            spec.act_dampratio[k] = float(aa.get("dampratio", "0"))
            if float(aa.get("gear", "1").split()[0]) != 1.0:
                raise ValueError("actuator gear != 1 is not supported")
            spec.act_ctrlrange[k] = _floats(aa["ctrlrange"]) if aa.get("ctrlrange") else (-np.inf, np.inf)
            spec.act_forcerange[k] = _floats(aa["forcerange"]) if aa.get("forcerange") else (-np.inf, np.inf)
    if seen != set(range(NJ)):
        raise ValueError("every arm joint needs one <position> actuator")
    _read_block(root, spec)
    return spec


def _geom_par(ga: dict, name: str, default) -> np.ndarray:
    v = _floats(ga[name]) if ga.get(name) else []
    return np.array(v + list(default[len(v):]), dtype=np.float64)


def _read_pads(pad_geoms, defaults: "_Defaults", spec: ModelSpec) -> None:
    """The arm's primitive box colliders (`finger_collision` pads on the jaws).  Mesh colliders are skipped: the STL
    files are not in the checkout.  All pads must share one set of contact parameters."""
    if not pad_geoms:
        return
    if len(pad_geoms) > MAX_PAD:
        raise ValueError(f"more than {MAX_PAD} primitive colliders on the arm")
    pos, size, par = [], [], None
    for body, ga in pad_geoms:
        if ga.get("quat") or ga.get("euler") or ga.get("axisangle") or ga.get("zaxis") or ga.get("xyaxes"):
            raise ValueError("rotated box colliders are not supported")
        if int(ga.get("condim", "3")) != 3 or int(ga.get("priority", "0")) != 0 or float(ga.get("solmix", "1")) != 1.0 \
                or float(ga.get("margin", "0")) != 0.0 or float(ga.get("gap", "0")) != 0.0:
            raise ValueError("box colliders must use condim 3 and default priority / solmix / margin / gap")
        pos.append(_floats(ga.get("pos", "0 0 0")))
        size.append(_floats(ga["size"]))
        p = (tuple(_geom_par(ga, "solref", _DEF_SOLREF)), tuple(_geom_par(ga, "solimp", _DEF_SOLIMP)),
             float(_geom_par(ga, "friction", (1.0, 0.005, 0.0001))[0]))
        if par is not None and p != par:
            raise ValueError("box colliders with different contact parameters are not supported")
        par = p
        spec.pad_body.append(body)
        spec.pad_names.append(ga.get("name", ""))
    spec.pad_pos, spec.pad_size = np.array(pos), np.array(size)
    spec.pad_solref, spec.pad_solimp, spec.pad_friction = np.array(par[0]), np.array(par[1]), par[2]


def _read_block(root, spec: ModelSpec) -> None:
    """The free block and the floor plane of the SCENE file (env01.xml:29-39): box half-size, mass, contact parameters.

    MuJoCo rules applied: `inertiafromgeom="true"` makes the box geom (default density 1000) define the mass and
    overrides the `<inertial>`; a contact pair takes the larger friction and, for equal priority and solmix, the
    mean of solref / solimp; the pair collides iff (contype1 & conaffinity2) | (contype2 & conaffinity1)."""
    comp = root.find("compiler")
    from_geom = comp is not None and comp.get("inertiafromgeom", "auto") == "true"
    blk = None
    for wb in root.findall("worldbody"):
        found = _find_body(wb, "block_a")
        blk = found if found is not None else blk
    floor = None
    for wb in root.findall("worldbody"):
        for g in wb.findall("geom"):
            if g.get("type") == "plane":
                floor = g
    if floor is not None:
        if any(abs(v) > 0 for v in _floats(floor.get("pos", "0 0 0"))) or floor.get("quat") or floor.get("euler"):
            raise ValueError("only the z = 0 floor plane is supported")
        spec.floor_solref = _geom_par(floor.attrib, "solref", _DEF_SOLREF)
        spec.floor_solimp = _geom_par(floor.attrib, "solimp", _DEF_SOLIMP)
        spec.floor_friction = float(_geom_par(floor.attrib, "friction", (1.0, 0.005, 0.0001))[0])
    else:  # no floor: nothing for the pads to touch
        spec.pad_body, spec.pad_names, spec.pad_pos, spec.pad_size = [], [], np.zeros((0, 3)), np.zeros((0, 3))
    if blk is None or floor is None:
        spec.block_ncon = 0
        return
    geom = blk.find("geom")
    if geom is None or geom.get("type") != "box":
        raise ValueError("block_a needs one box geom")
    size = _floats(geom.get("size"))
    if any(abs(v) > 0 for v in _floats(geom.get("pos", "0 0 0"))) or any(abs(v) > 0 for v in _floats(floor.get("pos", "0 0 0"))):
        raise ValueError("block geom / floor offsets are not supported")
    spec.block_half_z = size[2]
    inert = blk.find("inertial")
    if from_geom or inert is None:
        spec.block_mass = float(geom.get("density", "1000")) * 8.0 * size[0] * size[1] * size[2]
    else:
        spec.block_mass = float(inert.get("mass"))

    def par(g, name, default):
        v = _floats(g.get(name)) if g.get(name) else []
        return np.array(v + list(default[len(v):]))

    fr = max(par(geom, "friction", (1.0, 0.005, 0.0001))[0], par(floor, "friction", (1.0, 0.005, 0.0001))[0])
    spec.block_friction = float(fr)
    spec.contact_solref = 0.5 * (par(geom, "solref", _DEF_SOLREF) + par(floor, "solref", _DEF_SOLREF))
    spec.contact_solimp = 0.5 * (par(geom, "solimp", _DEF_SOLIMP) + par(floor, "solimp", _DEF_SOLIMP))
    ct = lambda g, k: int(g.get(k, "1"))  # noqa: E731
    collide = (ct(geom, "contype") & ct(floor, "conaffinity")) | (ct(floor, "contype") & ct(geom, "conaffinity"))
    if int(geom.get("condim", "3")) != 3 or int(floor.get("condim", "3")) != 3:
        raise ValueError("only condim 3 block / floor contacts are supported")
    spec.block_ncon = 4 if collide else 0


def reference_scene_path() -> str | None:
    """The reference's own MJCF, when the read-only checkout is mounted (never on the GPU box)."""
    p = "/root/reference/src/so100_mujoco_rl/envs/model/env01.xml"
    return p if os.path.exists(p) else None
