"""RMS pruning of conformers on the GPU (RDKit ``EmbedParameters.pruneRmsThresh``; reference:
``rdkit_extensions/conformer_pruning.cpp:96-137``, host-only there). One CTA per molecule walks its conformers in order and
keeps one only if its best-alignment RMSD to every conformer kept before it is at least the threshold
(``b200mol_rms_prune``)."""

from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from nvmolkit_b200 import _lib
from nvmolkit_b200._interop import require_cuda, stream_ctx, stream_ptr


def match_tables(matches: Sequence[Optional[np.ndarray]], atom_counts: Sequence[int]):
    """Per-molecule self-match tables ([K, L] index arrays; None = all atoms, identity) -> (offset [nMols+1] int32,
    length [nMols] int32, atoms int16) in the layout of b200mol_rms_prune."""
    offs, lens, flat = [0], [], []
    for m, n in zip(matches, atom_counts):
        a = np.arange(int(n), dtype=np.int16).reshape(1, -1) if m is None else np.ascontiguousarray(m, dtype=np.int16).reshape(len(m), -1)
        lens.append(a.shape[1])
        flat.append(a.ravel())
        offs.append(offs[-1] + a.size)
    return (np.array(offs, dtype=np.int32), np.array(lens, dtype=np.int32),
            np.concatenate(flat) if flat else np.zeros(0, np.int16))


def rms_prune(xyz: torch.Tensor, conf_atom_start: np.ndarray, mol_conf_start: np.ndarray, rms_thresh: float,
              matches: Optional[Sequence[Optional[np.ndarray]]] = None, atom_counts: Optional[Sequence[int]] = None,
              valid: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
    """uint8 keep flags [nConf] (device). `xyz` float64 [atoms, 3] on the GPU; conformers of molecule m are
    [mol_conf_start[m], mol_conf_start[m+1]) in embedding order; `matches` as in :func:`match_tables`."""
    sptr = stream_ptr(stream)
    require_cuda()
    if rms_thresh < 0:
        raise ValueError("rms_thresh must be >= 0")
    dev = xyz.device
    n_mols, n_conf = len(mol_conf_start) - 1, len(conf_atom_start) - 1
    with stream_ctx(stream):
        d_mcs = torch.from_numpy(np.ascontiguousarray(mol_conf_start, dtype=np.int32)).to(dev)
        d_cas = torch.from_numpy(np.ascontiguousarray(conf_atom_start, dtype=np.int32)).to(dev)
        keep = torch.zeros(n_conf, dtype=torch.uint8, device=dev)
        mo = ml = ma = None
        if matches is not None:
            off, ln, at = match_tables(matches, atom_counts)
            mo, ml, ma = (torch.from_numpy(a).to(dev) for a in (off, ln, at))
        v = valid.to(torch.uint8).contiguous() if valid is not None else None
        _lib.call("b200mol_rms_prune", n_mols, d_mcs.data_ptr(), d_cas.data_ptr(), xyz.contiguous().data_ptr(),
                  mo.data_ptr() if mo is not None else None, ml.data_ptr() if ml is not None else None,
                  ma.data_ptr() if ma is not None else None, float(rms_thresh), v.data_ptr() if v is not None else None,
                  keep.data_ptr(), sptr)
    return keep
