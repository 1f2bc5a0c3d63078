"""Drop-in mirror of the reference's ``kt_service`` package for the imaging hot path: same module
paths, class / function names, signatures, config files and log-and-sentinel error behaviour
(kt_service/ai_tools/ai_tools.py, utils.py, mesh_tools/femm_generator.py), with the arithmetic
running in libeitb200 on the GPU.  Parts the reference delegates to packages that are out of the
hot path (zip/DICOM decode, Gmsh meshing, pyEIT simulation, PNG collage) are explicit seams."""
