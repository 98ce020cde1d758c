"""Restart checkpoints in jQMC's HDF5 layout (format version 1.0), for the replacement drivers.

What the reference's CLI does around a run (jqmc/jqmc_cli.py:587-639, 692-743) is: every rank ``save_to_hdf5``s a temporary
per-rank file, rank 0 merges them with the Hamiltonian into ``restart.h5``, and a later job ``load_from_hdf5``s its rank.
This module provides the same four operations on the same file layout (jqmc/_checkpoint.py:1-30, 146-240, 330-400):

    restart.h5
    |-- _meta/                  attrs: format_version, driver_type, mpi_size, jqmc_version, timestamp
    |-- hamiltonian_data/       the dataclass tree, one group per dataclass, `_class_name` / `_module_name` attributes
    `-- rank_{R}/               driver_config (attrs), rng_state, walker_state, observables (datasets)

Dataclass trees follow ``_save_dataclass_to_hdf5`` / ``_load_dataclass_from_hdf5`` (jqmc/hamiltonians.py:369-573): numpy
arrays and homogeneous scalar sequences are datasets, nested dataclasses are groups, scalars are attributes, ``None`` is
absent.  ``_module_name`` carries the REFERENCE's module of each class, so that a file written here loads into jQMC's own
flax dataclasses, and a file written by jQMC loads into ``jqmc_b200.data``.

I/O goes through ``h5py`` when it is installed and through ``jqmc_b200.hdf5_lite`` otherwise (same bytes-level format family:
superblock v0, contiguous datasets; strings fixed-length UTF-8, booleans int8).
"""

from __future__ import annotations

import dataclasses
import datetime
import os
from typing import Any

import numpy as np

from . import data as D
from .hdf5_lite import open_file

CHECKPOINT_FORMAT_VERSION = "1.0"  # jqmc/_checkpoint.py:83

# module of each L1 dataclass in the reference (the loader of jqmc/hamiltonians.py:455-460 imports `_module_name`)
REFERENCE_MODULE = {
    "Hamiltonian_data": "jqmc.hamiltonians",
    "Structure_data": "jqmc.structure",
    "Coulomb_potential_data": "jqmc.coulomb_potential",
    "Wavefunction_data": "jqmc.wavefunction",
    "Jastrow_data": "jqmc.jastrow_factor",
    "Jastrow_one_body_data": "jqmc.jastrow_factor",
    "Jastrow_two_body_data": "jqmc.jastrow_factor",
    "Jastrow_three_body_data": "jqmc.jastrow_factor",
    "Geminal_data": "jqmc.determinant",
    "MOs_data": "jqmc.molecular_orbital",
    "AOs_sphe_data": "jqmc.atomic_orbital",
    "AOs_cart_data": "jqmc.atomic_orbital",
}
_SCALARS = (int, float, bool, str, np.number, np.bool_)


def _is_group(item) -> bool:
    return hasattr(item, "keys")


# ---- dataclass tree <-> HDF5 group ---------------------------------------------------------------------------------------
def _save_item(group, name: str, value: Any) -> None:
    if value is None:
        return
    if hasattr(value, "__dlpack__") and not isinstance(value, np.ndarray):  # jax / torch arrays
        value = np.asarray(value)
    if isinstance(value, np.ndarray):
        group.create_dataset(name, data=value)
    elif dataclasses.is_dataclass(value):
        save_dataclass_to_hdf5(group.create_group(name), value)
    elif isinstance(value, dict):
        sub = group.create_group(name)
        sub.attrs["_is_dict"] = True
        for k, v in value.items():
            _save_item(sub, str(k), v)
    elif isinstance(value, (list, tuple)):
        if len(value) == 0:
            group.create_dataset(name, data=np.array([]))
        elif all(isinstance(v, _SCALARS) for v in value):
            group.create_dataset(name, data=np.asarray(value))
        else:
            sub = group.create_group(name)
            sub.attrs["_is_list"] = True
            for i, v in enumerate(value):
                _save_item(sub, str(i), v)
    elif isinstance(value, _SCALARS):
        group.attrs[name] = value.item() if isinstance(value, (np.number, np.bool_)) else value


def save_dataclass_to_hdf5(group, obj) -> None:
    if not dataclasses.is_dataclass(obj):
        raise ValueError(f"{obj!r} is not a dataclass")
    cls = obj.__class__.__name__
    group.attrs["_class_name"] = cls
    group.attrs["_module_name"] = REFERENCE_MODULE.get(cls, obj.__class__.__module__)
    for f in dataclasses.fields(obj):
        _save_item(group, f.name, getattr(obj, f.name))


def _decode(v):
    if isinstance(v, bytes):
        return v.decode("utf-8")
    if isinstance(v, np.ndarray) and v.dtype.kind == "S":
        return np.char.decode(v, "utf-8")
    if isinstance(v, np.ndarray) and v.dtype.kind == "O" and v.size and isinstance(v.flat[0], bytes):
        return np.array([x.decode("utf-8") for x in v.flat]).reshape(v.shape)
    return v


def _load_item(item):
    if not _is_group(item):
        return _decode(item[()])
    attrs = item.attrs
    if attrs.get("_is_list"):
        return [_load_item(item[k]) for k in sorted(item.keys(), key=int)]
    if attrs.get("_is_dict"):
        return {k: _load_item(item[k]) for k in item.keys()}
    cls_name = _decode(attrs.get("_class_name"))
    if cls_name and hasattr(D, cls_name):
        return load_dataclass_from_hdf5(getattr(D, cls_name), item)
    if cls_name:
        raise NotImplementedError(f"{cls_name} has no counterpart in jqmc_b200.data (NN Jastrow and PBC are outside the engine)")
    return {k: _load_item(item[k]) for k in item.keys()}


def load_dataclass_from_hdf5(cls, group):
    kw = {}
    for f in dataclasses.fields(cls):
        ann = str(f.type)
        if f.name in group:
            val = _load_item(group[f.name])
            if isinstance(val, np.ndarray) and ("Sequence" in ann or "tuple" in ann or "list" in ann) and "ndarray" not in ann:
                val = tuple(val.tolist())  # the reference stores these fields as tuples (static pytree leaves)
            elif isinstance(val, list) and "Sequence" in ann:
                val = tuple(val)
            kw[f.name] = val
        elif f.name in group.attrs:
            val = _decode(group.attrs[f.name])
            if ann == "bool":
                val = bool(val)
            elif ann == "int":
                val = int(val)
            elif ann == "float":
                val = float(val)
            kw[f.name] = val
    return cls(**kw)


# ---- per-rank files, merge, load (jqmc/_checkpoint.py:146-400) -------------------------------------------------------------
def _put(group, mapping, scalars_as_attrs=True, skip_empty=False) -> None:
    for k, v in mapping.items():
        if v is None:
            continue
        if isinstance(v, dict):
            sub = group.create_group(k)
            for sk, sv in v.items():
                if isinstance(sv, np.ndarray) and sv.size > 0:
                    sub.create_dataset(sk, data=sv)
        elif isinstance(v, np.ndarray):
            if not (skip_empty and v.size == 0):
                group.create_dataset(k, data=v)
        elif scalars_as_attrs and isinstance(v, _SCALARS):
            group.attrs[k] = v.item() if isinstance(v, (np.number, np.bool_)) else v


def save_rank_checkpoint(filepath: str, *, driver_type: str, driver_config: dict, rng_state: dict, walker_state: dict,
                         observables: dict) -> None:  # fmt: skip
    """One rank's state as a stand-alone file with the four groups of a `rank_R` group (merged later by rank 0)."""
    with open_file(filepath, "w") as f:
        _put(f.create_group("driver_config"), driver_config)
        _put(f.create_group("rng_state"), rng_state)
        _put(f.create_group("walker_state"), walker_state)
        _put(f.create_group("observables"), observables, scalars_as_attrs=False, skip_empty=True)


def _copy_tree(src, dst) -> None:
    for k, v in src.attrs.items():
        dst.attrs[k] = v
    for k in src.keys():
        item = src[k]
        if _is_group(item):
            _copy_tree(item, dst.create_group(k))
        else:
            v = item[()]
            if isinstance(v, list):  # hdf5_lite returns string datasets as lists
                v = np.array([s.encode("utf-8") for s in v])
            dst.create_dataset(k, data=v)


def merge_rank_checkpoints(output_path: str, *, mpi_size: int, driver_type: str, hamiltonian_data,
                           tmp_pattern: str = "._restart_rank{rank}.h5", cleanup: bool = True) -> None:  # fmt: skip
    """Rank 0: `_meta` + `hamiltonian_data` + one `rank_R` group per temporary file -> one checkpoint."""
    if os.path.exists(output_path):
        os.remove(output_path)
    with open_file(output_path, "w") as out:
        meta = out.create_group("_meta")
        meta.attrs["format_version"] = CHECKPOINT_FORMAT_VERSION
        meta.attrs["driver_type"] = driver_type
        meta.attrs["mpi_size"] = int(mpi_size)
        meta.attrs["jqmc_version"] = "jqmc_b200"
        meta.attrs["timestamp"] = datetime.datetime.now(datetime.timezone.utc).isoformat()
        save_dataclass_to_hdf5(out.create_group("hamiltonian_data"), hamiltonian_data)
        for rank in range(mpi_size):
            tmp = tmp_pattern.format(rank=rank)
            with open_file(tmp, "r") as t:
                _copy_tree(t, out.create_group(f"rank_{rank}"))
            if cleanup:
                os.remove(tmp)


def _native(v):
    v = _decode(v)
    if isinstance(v, np.generic):
        return v.item()
    return v


def load_rank_checkpoint(filepath: str, rank: int) -> dict:
    out = {}
    with open_file(filepath, "r") as f:
        grp = f[f"rank_{rank}"]
        for name in ("driver_config", "rng_state", "walker_state"):
            g = grp[name]
            d = {k: _native(v) for k, v in g.attrs.items()}
            for k in g.keys():
                d[k] = _decode(g[k][()])
            out[name] = d
        obs = {}
        og = grp["observables"]
        for k in og.keys():
            item = og[k]
            obs[k] = {sk: item[sk][()] for sk in item.keys()} if _is_group(item) else item[()]
        out["observables"] = obs
    return out


def load_hamiltonian_from_checkpoint(filepath: str):
    with open_file(filepath, "r") as f:
        return load_dataclass_from_hdf5(D.Hamiltonian_data, f["hamiltonian_data"])


def load_checkpoint_meta(filepath: str) -> dict:
    with open_file(filepath, "r") as f:
        return {k: _native(v) for k, v in f["_meta"].attrs.items()}


def check_checkpoint_version(filepath: str) -> None:
    v = load_checkpoint_meta(filepath).get("format_version", "unknown")
    if v != CHECKPOINT_FORMAT_VERSION:
        raise ValueError(f"Checkpoint format version mismatch: file has '{v}', this driver expects '{CHECKPOINT_FORMAT_VERSION}'.")


def save_checkpoint(driver, output_path: str = "restart.h5", tmp_pattern: str = "._restart_rank{rank}.h5") -> None:
    """The CLI's write flow in one call (jqmc/jqmc_cli.py:587-600): every rank writes its temporary file, barrier, rank 0
    merges with the Hamiltonian, barrier."""
    from .mcmc import _dist, _rank_size

    rank, world = _rank_size()
    driver.save_to_hdf5(tmp_pattern.format(rank=rank))
    d = _dist()
    if d is not None and world > 1:
        d.barrier()
    if rank == 0:
        merge_rank_checkpoints(output_path, mpi_size=world, driver_type=type(driver).__name__, hamiltonian_data=driver.hamiltonian_data,
                               tmp_pattern=tmp_pattern)  # fmt: skip
    if d is not None and world > 1:
        d.barrier()
