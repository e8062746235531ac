for f in _alt/libobs_*.so; do ORCA_B200_LIB=$PWD/$f python tools/obs_time.py; done
