"""Scenario XML and config.cfg readers (TEST PLUMBING: feeds the same scenario files to the CUDA path and the oracle).

Python mirror of the reference's include/io/scenario.h:21-188 and include/io/configuration.h:68-121 (the C++
mirrors live in include/lbm/io/): same element / attribute names, same defaults, same document-order semantics.
A scenario becomes the box list the C ABI takes (lbm_b200_set_boxes): one entry per <boundary> node,
(kind, velocity, density, (x0, xE, y0, yE, z0, zE)), applied in order, last writer wins.
"""
import os
import xml.etree.ElementTree as ET

from lbm_b200 import capi

_CONDITIONS = {
    "noslip": capi.NOSLIP, "movingwall": capi.MOVINGWALL, "freeslip": capi.FREESLIP, "outflow": capi.OUTFLOW,
    "inflow": capi.INFLOW, "pressure": capi.PRESSURE, "periodic": capi.PERIODIC,
}


class ScenarioError(ValueError):
    pass


def _require(node, name):
    if name not in node.attrib:
        raise ScenarioError('Missing attribute "%s" for node "%s"!' % (name, node.tag))
    return node.attrib[name]


def _condition(node):
    """io/scenario.h:28-88: kind + constructor arguments of one <boundary> node"""
    cond = _require(node, "condition")
    if cond not in _CONDITIONS:
        raise ScenarioError(cond + " boundary condition not supported!")
    kind = _CONDITIONS[cond]
    v, rho = (0.0, 0.0, 0.0), 1.0
    if cond in ("movingwall", "inflow"):
        v = tuple(float(_require(node, k)) for k in ("vx", "vy", "vz"))
    if cond in ("outflow", "inflow") and "rho-ref" in node.attrib:      # optional, default 1.0
        rho = float(node.attrib["rho-ref"])
    if cond == "pressure":
        rho = float(_require(node, "rho-in"))
    return kind, v, rho


def _extent(text, xl, yl, zl):
    """io/scenario.h:99-127: named faces span whole ghost planes (edges included); otherwise six inclusive indices"""
    faces = {
        "z0": (0, xl + 1, 0, yl + 1, 0, 0), "zmax": (0, xl + 1, 0, yl + 1, zl + 1, zl + 1),
        "x0": (0, 0, 0, yl + 1, 0, zl + 1), "xmax": (xl + 1, xl + 1, 0, yl + 1, 0, zl + 1),
        "y0": (0, xl + 1, 0, 0, 0, zl + 1), "ymax": (0, xl + 1, yl + 1, yl + 1, 0, zl + 1),
    }
    if text in faces:
        return faces[text]
    try:
        ex = [int(t) for t in text.split()]
        if any(e < 0 for e in ex):
            raise ValueError
    except ValueError:
        raise ScenarioError("bad lexical cast: source type value could not be interpreted as target")
    if len(ex) < 6:
        raise ScenarioError('Extent "%s" is not complete! Must be six values!' % text)
    return tuple(ex[:6])


def read_fluid_mask(path):
    """legacy-VTK ASCII STRUCTURED_POINTS mask (io/vtk.hpp:94-157): returns (xl, yl, zl, origin, spacing, mask)"""
    import numpy as np
    try:
        tokens = open(path).read().split("\n")
    except OSError:
        raise ScenarioError('VTK file "%s" does not exist or does not seem to be a valid structured grids file!' % path)
    if not tokens or not tokens[0].startswith("# vtk"):
        raise ScenarioError('VTK file "%s" does not exist or does not seem to be a valid structured grids file!' % path)
    words = " ".join(tokens[2:]).split()
    dims, origin, spacing, n_points, i = None, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), 0, 0
    structured = False
    while i < len(words):
        w = words[i]
        if w == "DATASET":
            structured = words[i + 1] == "STRUCTURED_POINTS"; i += 2
        elif w == "DIMENSIONS":
            dims = tuple(int(v) for v in words[i + 1:i + 4]); i += 4
        elif w == "ORIGIN":
            origin = tuple(float(v) for v in words[i + 1:i + 4]); i += 4
        elif w in ("SPACING", "ASPECT_RATIO"):
            spacing = tuple(float(v) for v in words[i + 1:i + 4]); i += 4
        elif w == "POINT_DATA":
            n_points = int(words[i + 1]); i += 2
        elif w == "LOOKUP_TABLE":
            i += 2
            break
        else:
            i += 1
    if not structured or dims is None:
        raise ScenarioError('VTK file "%s" does not exist or does not seem to be a valid structured grids file!' % path)
    n = dims[0] * dims[1] * dims[2]
    if n_points != n or len(words) - i < n:
        raise ScenarioError("Could not read file!")
    mask = (np.array(words[i:i + n], dtype=np.int64) != 0).astype(np.uint8).reshape(dims[2], dims[1], dims[0])
    return dims[0], dims[1], dims[2], origin, spacing, mask


def load_scenario(path):
    """dict(name, xl, yl, zl, boxes, fluid_mask, origin, spacing) from a scenario XML file"""
    try:
        root = ET.parse(path).getroot()
    except (ET.ParseError, OSError):
        raise ScenarioError('XML file "%s" could not be read properly!' % path)
    if root.tag != "scenario":
        raise ScenarioError("Scenario node missing!")
    if "name" not in root.attrib:
        raise ScenarioError("Scenario name is missing!")
    dom = root.find("domain")
    if dom is None:
        raise ScenarioError("Domain node is missing!")
    mask, origin, spacing = None, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    if "vtk-file" in dom.attrib:
        vtk = dom.attrib["vtk-file"]
        if not os.path.isabs(vtk) and not os.path.exists(vtk):
            vtk = os.path.join(os.path.dirname(os.path.abspath(path)), vtk)
        xl, yl, zl, origin, spacing, mask = read_fluid_mask(vtk)
    elif all(k in dom.attrib for k in ("xl", "yl", "zl")):
        xl, yl, zl = (int(dom.attrib[k]) for k in ("xl", "yl", "zl"))
        if not (xl > 0 and yl > 0 and zl > 0):
            raise ScenarioError("Domain lengths xl/yl/zl must be positive!")
    else:
        raise ScenarioError("Neither vtk-file nor xl/yl/zl attribute provided to domain node!")
    boxes = []
    for node in dom:
        if node.tag != "boundary":
            continue
        ext = _extent(_require(node, "extent"), xl, yl, zl)
        kind, v, rho = _condition(node)
        boxes.append((kind, v, rho, ext))
    return dict(name=root.attrib["name"], xl=xl, yl=yl, zl=zl, boxes=boxes, fluid_mask=mask, origin=origin, spacing=spacing)


def load_config(path):
    """key = value file of io/configuration.h:82-120 with the reference's defaults and required keys"""
    cfg = {"collision-model": "bgk", "output-dir": "vtk", "omp-threads": 1, "gpus": 1, "lattice": 19, "arithmetic": "fast"}
    known = set(cfg) | {"tau", "timesteps", "timesteps-per-plot", "scenario-file"}
    with open(path) as fh:
        for line in fh:
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            if "=" not in line:
                raise ScenarioError("the options configuration file contains an invalid line '%s'" % line)
            key, value = (t.strip() for t in line.split("=", 1))
            if key not in known:
                raise ScenarioError("unrecognised option '%s'" % key)
            cfg[key] = value
    for key in ("tau", "timesteps", "timesteps-per-plot", "scenario-file"):
        if key not in cfg:
            raise ScenarioError("the option '--%s' is required but missing" % key)
    cfg["tau"] = float(cfg["tau"])
    for key in ("timesteps", "timesteps-per-plot", "omp-threads", "gpus", "lattice"):
        cfg[key] = int(cfg[key])
    if not (0.5 < cfg["tau"] < 2.0):
        raise ScenarioError("tau must be in (0.5, 2.0)")
    if cfg["collision-model"] != "bgk":
        raise ScenarioError("only the bgk collision model is supported")
    return cfg


def domain_from_scenario(path, Q, tau, **kw):
    """capi.Domain with the scenario's mask and boundaries applied (io/scenario.h:131-188)"""
    sc = load_scenario(path)
    d = capi.Domain(Q, sc["xl"], sc["yl"], sc["zl"], tau, **kw)
    if sc["fluid_mask"] is not None:
        d.set_fluid_mask(sc["fluid_mask"])
    if sc["boxes"]:
        d.set_boxes(sc["boxes"])
    return d, sc
