"""The .Call shim (r_package/src/epx_rshim.c) compiled against the fake R API of tests/_rstub/: what can be checked
without a GPU -- it builds warning-free, registers the routines the R wrapper calls, and every malformed argument is an R
error raised BEFORE any device work, with the protect stack and the R_alloc stack unwound."""
import numpy as np

import rshim


def test_shim_registers_the_call_routines():
    routines, dynamic = rshim.registered()
    assert routines == {"C_epx_upload": 2, "C_epx_analyze": 5, "C_epx_groups": 3, "C_epx_release": 1,
                        "C_epx_device_info": 0}
    assert dynamic == 0          # R_useDynamicSymbols(dll, FALSE)
    # every routine the R wrapper names is registered
    src = open(rshim.os.path.join(rshim.HERE, "..", "r_package", "R", "epimethex_analysis.R")).read()
    import re
    assert set(re.findall(r"\.Call\((C_epx_\w+)", src)) <= set(routines)


def test_malformed_upload_arguments_are_r_errors_without_leaks():
    cols = [np.random.default_rng(j).random(10) for j in range(4)]
    for bad_kind, text in ((1, "rows"), (2, "not numeric"), (3, "device must be an integer vector")):
        h, rep = rshim.upload(cols, bad_kind=bad_kind)
        assert h is None and rep.rc == 1 and text in rep.message, rep.message
        assert rep.protect_depth == 0 and rep.ralloc_live == 0 and rep.live_finalizers == 0
    h, rep = rshim.upload(cols[:2])
    assert h is None and rep.rc == 1 and "at least 3" in rep.message
    assert rep.collections > 0   # the stub's collector ran (at every allocation)
    rshim.finish()
