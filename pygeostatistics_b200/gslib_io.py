"""Geo-EAS / gslib files and the JSON parameter files of the krige3d path.

Mirrors the reference's boundary:
  SpatialData        gslib_reader.py:17-50 (the reader; statistics/plots are out of scope)
  loader_patched     yaml_patch.py:14-26   (float resolver so `1e+21` parses as a float)
  write_gslib_grid   the output layout of gslib_help/kt3d_help.md:37-41, which the
                     reference parses a name for (krige3d.py:83 `outfl`) but never writes
"""
import json
import re

import numpy as np

try:  # same third-party parsers as the reference when present; plain numpy/json otherwise
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None
try:
    import yaml
except Exception:  # pragma: no cover
    yaml = None


def loader_patched():
    """yaml.SafeLoader with the scientific-notation float resolver (yaml_patch.py:14-26)."""
    loader = yaml.SafeLoader
    loader.add_implicit_resolver(
        u'tag:yaml.org,2002:float',
        re.compile(u'''^(?:
        [-+]?(?:[0-9][0-9_]*)\\.[0-9_]*(?:[eE][-+]?[0-9]+)?
        |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
        |\\.[0-9_]+(?:[eE][-+][0-9]+)?
        |[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+\\.[0-9_]*
        |[-+]?\\.(?:inf|Inf|INF)
        |\\.(?:nan|NaN|NAN))$''', re.X),
        list(u'-+0123456789.'))
    return loader


def read_params(path):
    """The .par files are JSON objects (parameters/krige3d_write_params.py:80-82)."""
    with open(path, "r") as fin:
        if yaml is not None:
            return yaml.load(fin, Loader=loader_patched())
        return json.load(fin)


class SpatialData(object):
    """Geo-EAS reader: title line, column count, column names, tab separated rows.
    `vr` is a numpy record array of f8 fields named by the header, with a zero `z`
    column appended when the file has none (gslib_reader.py:25-50)."""

    def __init__(self, file_path):
        self.datafl = file_path
        self.vr = None
        self.property_name = None
        self._2d = False
        self._read_data()

    def _read_data(self):
        column_name = []
        with open(self.datafl, 'r') as fin:
            fin.readline()
            ncols = int(fin.readline().strip())
            for _ in range(ncols):
                column_name.append(fin.readline().strip())
        self.property_name = [item for item in column_name if item not in ['x', 'y', 'z']]
        if pd is not None:
            values = pd.read_csv(self.datafl, sep='\t', header=None, names=column_name,
                                 skiprows=ncols + 2).values
        else:  # pragma: no cover
            values = np.loadtxt(self.datafl, delimiter='\t', skiprows=ncols + 2, ndmin=2)
        values = np.asarray(values, dtype=np.float64).reshape(-1, ncols)
        arrays = [np.ascontiguousarray(values[:, i]) for i in range(ncols)]
        if 'z' not in column_name:
            self._2d = True
            column_name.append('z')
            arrays.append(np.zeros(values.shape[0]))
        dtype = np.dtype({'names': column_name, 'formats': ['f8'] * len(column_name)})
        self.vr = np.rec.fromarrays(arrays, dtype=dtype)


def read_grid_column(path, col):
    """One column (0-based `col`, or a column name) of a Geo-EAS gridded file, x fastest:
    the external-drift file `secfl` (gslib_help/kt3d_help.md:93-100)."""
    with open(path, 'r') as fin:
        fin.readline()
        ncols = int(fin.readline().strip())
        names = [fin.readline().strip() for _ in range(ncols)]
    if isinstance(col, str):
        col = names.index(col)
    if col < 0 or col >= ncols:
        raise ValueError("column {} not in {}".format(col, path))
    if pd is not None:
        v = pd.read_csv(path, sep=r'\s+', header=None, skiprows=ncols + 2,
                        usecols=[col]).values[:, 0]
    else:  # pragma: no cover
        v = np.loadtxt(path, skiprows=ncols + 2, usecols=[col], ndmin=1)
    return np.ascontiguousarray(v, dtype=np.float64)


def write_gslib_grid(path, estimate, variance, title="kt3d estimates", unest=None):
    """Output grid, x fastest then y then z: two columns Estimate, EstimationVariance.
    `unest` replaces NaN (GSLIB convention -999.); None keeps NaN as the reference's
    in-memory arrays do (krige3d.py:189)."""
    est = np.asarray(estimate, dtype=np.float64)
    var = np.asarray(variance, dtype=np.float64)
    if unest is not None:
        est = np.where(np.isnan(est), unest, est)
        var = np.where(np.isnan(var), unest, var)
    with open(path, "w") as fout:
        fout.write("%s\n2\nEstimate\nEstimationVariance\n" % title)
        np.savetxt(fout, np.column_stack([est, var]), fmt="%.17g", delimiter="\t")


def write_gslib_validation(path, xyz, true, estimate, variance, title="kt3d cross validation",
                           unest=None):
    """Cross-validation / jackknife output, one record per location: X, Y, Z, True,
    Estimate, EstimationVariance, Error = estimate - true (GSLIB kt3d's layout for
    `option` 1 and 2; gslib_help/kt3d_help.md:20-27 names the option, the reference writes
    nothing)."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    cols = [xyz[:, 0], xyz[:, 1], xyz[:, 2]] + [np.asarray(a, dtype=np.float64)
                                                for a in (true, estimate, variance)]
    cols.append(cols[4] - cols[3])
    table = np.column_stack(cols)
    if unest is not None:
        table = np.where(np.isnan(table), unest, table)
    with open(path, "w") as fout:
        fout.write("%s\n7\nX\nY\nZ\nTrue\nEstimate\nEstimationVariance\nError: est-true\n" % title)
        np.savetxt(fout, table, fmt="%.17g", delimiter="\t")
