"""Drop-in `env` package for the reference's scripts (N = 1 facade over the batched CUDA backend).

Put this directory in front of the reference's `src/` on PYTHONPATH:

    cd /path/to/HOPE/src && PYTHONPATH=/path/to/repo/hope_b200/compat:/path/to/repo python train/train_HOPE_sac.py ...

`env.car_parking_base.CarParking`, `env.env_wrapper.CarParkingWrapper` and `env.vehicle.{Status,VALID_SPEED}`
then resolve to the modules here (the scripts append `..` and `.` to sys.path, so a PYTHONPATH entry
shadows their own `env`), while `configs`, `model.*` and `evaluation.*` stay the reference's.
"""
