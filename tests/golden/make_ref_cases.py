"""Copies the INPUT DATA (maps, tables, ini -- no source code) of the reference's own acceptance
cases to where the tests find them: baseline/_ref/, which is git-ignored (the repository holds no copy
of reference files) but travels to the GPU box with gpurun.  Run in the build container, where
/root/reference exists (__graft_entry__.build() does it when the data is missing):

    python tests/golden/make_ref_cases.py

* tests/wflow_sbm (the case of the reference's tests/test_wflow_sbm.py; forcing comes through the
  API, so only 8 static maps + the lookup tables + the ini are needed, ~1 MB; plus what its dynamic-vegetation
  variant, tests/test_wflow_sbm_dynveg.py, reads: 3 tables, the land-cover map, 2 months of LAI)
      -> baseline/_ref/ref_cases/wflow_sbm/
* tests/wflow_hbv (the case of the reference's tests/test_wflow_hbv.py; static maps + tables + ini)
      -> baseline/_ref/ref_cases/wflow_hbv/
* examples/wflow_rhine_sbm (the case of tests/test_wflow_sbm_example.py; 132 MB with its forcing)
      -> baseline/_ref/wflow_rhine_sbm/
"""
import os
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

MAPS = ["wflow_subcatch", "wflow_ldd", "wflow_dem", "wflow_river", "wflow_landuse", "wflow_soil", "wflow_gauges",
        "wflow_riverlength_fact"]


def main():
    src = os.path.join(REF, "tests", "wflow_sbm")
    dst = os.path.join(ROOT, "baseline", "_ref", "ref_cases", "wflow_sbm")
    os.makedirs(os.path.join(dst, "staticmaps"), exist_ok=True)
    for m in MAPS:
        shutil.copyfile(os.path.join(src, "staticmaps", m + ".map"), os.path.join(dst, "staticmaps", m + ".map"))
    shutil.copytree(os.path.join(src, "intbl"), os.path.join(dst, "intbl"), dirs_exist_ok=True)
    shutil.copyfile(os.path.join(src, "wflow_sbm.ini"), os.path.join(dst, "wflow_sbm.ini"))
    # the dynamic-vegetation variant of the same case (tests/test_wflow_sbm_dynveg.py): its ini, the land-cover
    # lookup tables and the LAI climatology of the two months the 29 steps touch
    shutil.copyfile(os.path.join(src, "wflow_sbm_dynveg.ini"), os.path.join(dst, "wflow_sbm_dynveg.ini"))
    os.makedirs(os.path.join(dst, "inmaps", "clim"), exist_ok=True)
    for f in ("LC.map", "LCtoBranchTrunkStorage.tbl", "LCtoExtinctionCoefficient.tbl", "LCtoSpecificLeafStorage.tbl",
              "LAI00000.001", "LAI00000.002"):
        shutil.copyfile(os.path.join(src, "inmaps", "clim", f), os.path.join(dst, "inmaps", "clim", f))
    # tests/wflow_hbv (the case of the reference's tests/test_wflow_hbv.py: cold start, forcing through the API)
    hsrc = os.path.join(REF, "tests", "wflow_hbv")
    hdst = os.path.join(ROOT, "baseline", "_ref", "ref_cases", "wflow_hbv")
    os.makedirs(os.path.join(hdst, "staticmaps"), exist_ok=True)
    for m in MAPS:
        shutil.copyfile(os.path.join(hsrc, "staticmaps", m + ".map"), os.path.join(hdst, "staticmaps", m + ".map"))
    shutil.copytree(os.path.join(hsrc, "intbl"), os.path.join(hdst, "intbl"), dirs_exist_ok=True)
    shutil.copyfile(os.path.join(hsrc, "wflow_hbv.ini"), os.path.join(hdst, "wflow_hbv.ini"))
    shutil.copyfile(os.path.join(hsrc, "wflow_hbv_hr.ini"), os.path.join(hdst, "wflow_hbv_hr.ini"))  # tests/test_wflow_hbv2.py
    os.system("chmod -R u+w '%s'" % hdst)
    # examples/wflow_rhine_hbv as shipped (warm start from instate, stream-order roughness, Courant sub-steps): static maps,
    # tables, states and the forcing of the first ten days (the shipped run is five days long)
    esrc = os.path.join(REF, "examples", "wflow_rhine_hbv")
    edst = os.path.join(ROOT, "baseline", "_ref", "wflow_rhine_hbv")
    for sub in ("intbl", "instate"):
        shutil.copytree(os.path.join(esrc, sub), os.path.join(edst, sub), dirs_exist_ok=True)
    os.makedirs(os.path.join(edst, "staticmaps"), exist_ok=True)
    for m in MAPS + ["wflow_streamorder"]:
        shutil.copyfile(os.path.join(esrc, "staticmaps", m + ".map"), os.path.join(edst, "staticmaps", m + ".map"))
    os.makedirs(os.path.join(edst, "inmaps"), exist_ok=True)
    for stem in ("P0000000", "PET00000", "TEMP0000"):
        for t in range(1, 11):
            f = "%s.%03d" % (stem, t)
            shutil.copyfile(os.path.join(esrc, "inmaps", f), os.path.join(edst, "inmaps", f))
    shutil.copyfile(os.path.join(esrc, "wflow_hbv.ini"), os.path.join(edst, "wflow_hbv.ini"))
    os.system("chmod -R u+w '%s'" % edst)
    rh = os.path.join(ROOT, "baseline", "_ref", "wflow_rhine_sbm")
    if not os.path.isdir(rh):
        shutil.copytree(os.path.join(REF, "examples", "wflow_rhine_sbm"), rh)
    for d in (dst, rh):
        os.system("chmod -R u+w '%s'" % d)
    print("wrote", dst, "and", rh)


if __name__ == "__main__":
    main()
