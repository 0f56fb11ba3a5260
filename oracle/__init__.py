"""TEST INFRASTRUCTURE ONLY.

CPU restatement (``gofrt_oracle.c``) of the reference's g(r,t) path plus, when it has been built,
the unmodified reference itself (``oracle/_ref/analisi_ref*.so``, see ``oracle/Makefile``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package.  The product package ``analisi_b200`` never does.
"""
from .oracle import (  # noqa: F401
    OracleError,
    build,
    counts,
    vdata,
    mediavar,
    neighbour_hist,
    msd,
    cm_positions,
    min_image,
    d2_all,
    pbc_wrap,
    type_ids,
    itype,
    nextra,
    leff,
    lammps_to_internal,
    internal_to_lammps,
    load_ref,
)
