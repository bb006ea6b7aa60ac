"""Load the UNMODIFIED reference classes from /root/reference for pinning the oracle.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this module.
/root/reference does not exist on the GPU box, so this loader is only used
 (a) by oracle/make_golden.py, run in the build container to emit tests/golden/*.npz
 (b) by `-m "not gpu"` tests, which skip when the tree is absent.

The reference is Python 2.7 + ROS 1.  To execute it under
Python 3 / NumPy 2 without editing it we
  * register empty stand-ins for rospy / tf / *_msgs in sys.modules,
  * restore the removed alias np.int,
  * for icp.py only: rewrite py2 `print x` statements into calls (pure syntax,
    no arithmetic is touched) and exec the result into a fresh module.
bresenham.py and mapping.py import unchanged.
"""
import importlib.util
import math
import os
import re
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("B2S_REFERENCE_ROOT", "/root/reference")

W9_SCRIPTS = os.path.join(
    REF_ROOT, "W9_Fusion Localization (LiDAR Odometry)", "course_agv_slam", "scripts")
W12_MAPPING = os.path.join(
    REF_ROOT, "W12_LiDAR SLAM", "w12-mapping", "course_agv_slam", "scripts")
W12_ONLINE = os.path.join(
    REF_ROOT, "W12_LiDAR SLAM", "w12-mapping-online", "course_agv_slam", "scripts")
W12_FINAL = os.path.join(
    REF_ROOT, "W12_LiDAR SLAM", "w12-ekf-slam-final", "course_agv_slam", "scripts")


def available():
    return os.path.isfile(os.path.join(W9_SCRIPTS, "icp.py"))


class _Anything(object):
    """Callable attribute sink used for ROS publishers / broadcasters."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


_ros_params = {}


def set_ros_params(params):
    """Values returned by the stubbed rospy.get_param (e.g. {'/icp/tolerance': 0})."""
    _ros_params.clear()
    _ros_params.update(params or {})


def _install_ros_stubs():
    if "rospy" in sys.modules and getattr(sys.modules["rospy"], "_b2s_stub", False):
        return

    def get_param(key, default=None):
        return _ros_params.get(key, default)

    rospy = types.ModuleType("rospy")
    rospy._b2s_stub = True
    rospy.get_param = get_param
    rospy.Publisher = _Anything
    rospy.Subscriber = _Anything
    rospy.Time = _Anything()          # an instance: rospy.Time.now() is called on the class in the reference
    rospy.init_node = lambda *a, **k: None
    rospy.spin = lambda *a, **k: None
    sys.modules["rospy"] = rospy

    tf = types.ModuleType("tf")
    tf.TransformBroadcaster = _Anything
    tf.transformations = types.SimpleNamespace(
        # yaw-only quaternion (x, y, z, w): all publishResult needs from tf ([ICP]:195); not on the measured path
        quaternion_from_euler=lambda roll, pitch, yaw: (0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0)))
    sys.modules["tf"] = tf

    for pkg, names in (
        ("sensor_msgs", ["LaserScan"]),
        ("nav_msgs", ["Odometry", "OccupancyGrid"]),
        ("geometry_msgs", ["TransformStamped"]),
        ("visualization_msgs", ["MarkerArray", "Marker"]),
    ):
        top = types.ModuleType(pkg)
        msg = types.ModuleType(pkg + ".msg")
        for n in names:
            setattr(msg, n, _Anything)
        top.msg = msg
        sys.modules[pkg] = top
        sys.modules[pkg + ".msg"] = msg
    srv = types.ModuleType("nav_msgs.srv")
    srv.GetMap = _Anything
    sys.modules["nav_msgs"].srv = srv
    sys.modules["nav_msgs.srv"] = srv

    if not hasattr(np, "int"):
        np.int = int  # removed in NumPy 1.24; the reference uses dtype=np.int


_PRINT_STMT = re.compile(r"^(\s*)print\s+(?!\()(.*)$")
_PRINT_BARE = re.compile(r"^(\s*)print\s*$")


def _py2_prints_to_calls(source):
    out = []
    for line in source.splitlines():
        m = _PRINT_STMT.match(line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        else:
            m = _PRINT_BARE.match(line)
            if m:
                line = "%sprint()" % m.group(1)
        out.append(line)
    return "\n".join(out) + "\n"


def _exec_module(name, path, rewrite_prints):
    with open(path, "r") as fh:
        src = fh.read()
    if rewrite_prints:
        src = _py2_prints_to_calls(src)
    mod = types.ModuleType(name)
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load_icp_class(params=None, silent=True):
    """The canonical ICP class, W9 icp.py (identical to the three W12 copies)."""
    _install_ros_stubs()
    set_ros_params(params)
    mod = _exec_module("_ref_icp", os.path.join(W9_SCRIPTS, "icp.py"), True)
    if silent:
        mod.__dict__["print"] = lambda *a, **k: None
    return mod.ICP


def load_fhb_icp_class():
    """W12 icp-fhb.py: pure NumPy peer copy (second, shim-free witness)."""
    _install_ros_stubs()
    path = os.path.join(W12_FINAL, "icp-fhb.py")
    spec = importlib.util.spec_from_file_location("_ref_icp_fhb", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ICP


def _load_mapping_from(scripts_dir, tag):
    _install_ros_stubs()
    bres = _exec_module("bresenham", os.path.join(scripts_dir, "bresenham.py"), False)
    saved = sys.modules.get("bresenham")
    sys.modules["bresenham"] = bres  # mapping.py does `import bresenham as drawing`
    try:
        mp = _exec_module("_ref_mapping_" + tag, os.path.join(scripts_dir, "mapping.py"), False)
    finally:
        if saved is None:
            sys.modules.pop("bresenham", None)
        else:
            sys.modules["bresenham"] = saved
    return mp.Mapping, bres.bresenham


def load_mapping_classes():
    """(Mapping, bresenham) from w12-mapping: endpoint weight +20."""
    return _load_mapping_from(W12_MAPPING, "w20")


def load_mapping_online_classes():
    """(Mapping, bresenham) from w12-mapping-online: endpoint weight +4."""
    return _load_mapping_from(W12_ONLINE, "w4")


def load_slam_node_class():
    """SLAM_EKF from w12-mapping slam_ekf.py, for its laserToNumpy / u2T / T2u helpers (the steps
    around the hot path).  Sibling modules it imports are stubbed except `mapping` (the real one)."""
    _install_ros_stubs()
    Mapping, _ = load_mapping_classes()
    stubs = {}
    for name, attrs in (("icp", {"ICP": _Anything}),
                        ("ekf_lm", {"EKF": _Anything, "STATE_SIZE": 3}),
                        ("extraction", {"LandMarkSet": _Anything, "Extraction": _Anything}),
                        ("mapping", {"Mapping": Mapping})):
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        stubs[name] = mod
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        node = _exec_module("_ref_slam_ekf", os.path.join(W12_MAPPING, "slam_ekf.py"), True)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    node.__dict__["print"] = lambda *a, **k: None
    return node.SLAM_EKF, Mapping


def load_localization_class():
    """Localization from W9 localization.py, for laserEstimation (the virtual scan of the static map, :128-150) --
    called as an unbound function on a duck-typed object.  Its sibling imports are stubbed except `icp` (the real one)."""
    _install_ros_stubs()
    ICP = load_icp_class({})
    stubs = {}
    for name, attrs in (("icp", {"ICP": ICP}), ("ekf", {"EKF": _Anything})):
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        stubs[name] = mod
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        node = _exec_module("_ref_localization", os.path.join(W9_SCRIPTS, "localization.py"), True)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    node.__dict__["print"] = lambda *a, **k: None
    return node.Localization
