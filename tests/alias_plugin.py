"""pytest plugin (``-p tests.alias_plugin``): mounts helicon_b200 at the reference's module paths, the way INTEGRATION.md
describes the drop-in, so that the reference's OWN test files (tests/test_denovo3D_solver.py:5,
tests/test_denovo3D_pipeline.py:5 import ``helicon.webApps.denovo3D.*``) exercise the CUDA path unchanged."""
import sys
import types


def _install():
    from helicon_b200 import pipeline, planner, solver_linear_regression

    helicon = types.ModuleType("helicon")
    helicon.__path__ = []
    webapps = types.ModuleType("helicon.webApps")
    webapps.__path__ = []
    denovo = types.ModuleType("helicon.webApps.denovo3D")
    denovo.__path__ = []
    helicon.webApps, webapps.denovo3D = webapps, denovo
    denovo.pipeline, denovo.solver_linear_regression = pipeline, solver_linear_regression
    # names of the helicon namespace the two test files touch (lib/io_mrc.py:71-98, lib/analysis.py:752-799)
    helicon.read_image_2d = pipeline.read_image_2d
    helicon.get_cylindrical_mask = planner.get_cylindrical_mask
    sys.modules.update({"helicon": helicon, "helicon.webApps": webapps, "helicon.webApps.denovo3D": denovo,
                        "helicon.webApps.denovo3D.pipeline": pipeline,
                        "helicon.webApps.denovo3D.solver_linear_regression": solver_linear_regression})
    if "mrcfile" not in sys.modules:  # the tests patch mrcfile.open; the package itself is optional here
        try:
            import mrcfile  # noqa: F401
        except ImportError:
            m = types.ModuleType("mrcfile")
            m.open = lambda *a, **k: (_ for _ in ()).throw(IOError("mrcfile is not installed"))
            sys.modules["mrcfile"] = m


_install()
