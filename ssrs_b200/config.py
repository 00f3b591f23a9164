"""`Config` — the reference's run configuration (API contract of `/root/reference/ssrs/config.py:9-67`).

Field names, order, defaults and the grouped `__str__` layout match the reference so user scripts
(`Config(...)`, `dataclasses.replace(cfg, ...)`, `Simulator(cfg)`) keep working unchanged.  B200-specific
knobs are *not* added here; they are keyword-only arguments of `ssrs_b200.Simulator`.
"""
import os
from dataclasses import dataclass, fields
from typing import Tuple

# (section title, first field of the section) in declaration order — drives __str__
_SECTIONS = (
    ("General settings", "run_name"),
    ("Terrain settings", "southwest_lonlat"),
    ("Uniform mode", "uniform_winddirn"),
    ("Snapshot mode", "snapshot_datetime"),
    ("Seasonal mode", "seasonal_start"),
    ("WindToolKit settings", "wtk_source"),
    ("Updraft computation", "thermals_realization_count"),
    ("Simulating tracks", "movement_model"),     # the reference breaks by position: index 23
    ("Plotting and wind turbines", "turbine_minimum_hubheight"),
)


@dataclass
class Config:
    """Configuration parameters for SSRS simulation """

    # -- general
    run_name: str = 'default'
    out_dir: str = os.path.join(os.path.abspath(os.path.curdir), 'output')
    max_cores: int = 8                     # kept for compatibility; the GPU path has no process pool
    sim_seed: int = -1                     # >= 0: seeds numpy (starting cells) and keys the stepper's Philox streams
    sim_mode: str = 'uniform'              # 'uniform' | 'snapshot' | 'seasonal'
    print_verbose: bool = False

    # -- terrain
    southwest_lonlat: Tuple[float, float] = (-106.21, 42.78)
    projected_crs: str = 'ESRI:102008'
    region_width_km: Tuple[float, float] = (60., 50.)     # with `resolution`: gridsize = (rows, cols) of every raster
    resolution: int = 100.                 # metres per cell (stage 1 stencil spacing; 10 m -> 5000 x 6000 cells)

    # -- uniform mode (direction: northerly 0, easterly 90, westerly 270)
    uniform_winddirn: float = 270.         # scalar wind of `ssrs_updraft` in uniform mode
    uniform_windspeed: float = 10.

    # -- snapshot mode
    snapshot_datetime: Tuple[int, int, int, int] = (2010, 6, 17, 13)

    # -- seasonal mode
    seasonal_start: Tuple[int, int] = (3, 20)
    seasonal_end: Tuple[int, int] = (5, 15)
    seasonal_timeofday: str = 'daytime'    # morning | afternoon | evening | daytime
    seasonal_count: int = 8

    # -- WIND Toolkit
    wtk_source: str = 'AWS'
    wtk_orographic_height: int = 100
    wtk_thermal_height: int = 100
    wtk_interp_type: str = 'linear'        # site -> grid: 'linear' (ssrs_interp_wind) | 'nearest' (ssrs_interp_wind_nearest)

    # -- updraft
    thermals_realization_count: bool = 0   # > 0: that many extra (updraft, potential, tracks) triples per wind case
    updraft_threshold: float = 0.75        # `thr` of get_above_threshold_speed, fused into the stencil kernel
    movement_model: str = 'fluidflow'      # 'fluidflow' (potential solve + field-driven steps) | 'drw' (no fields)

    # -- tracks
    track_direction: float = 0             # degrees; picks the Dirichlet sets of the solve and the directional weights
    track_count: str = 1000                # tracks per (wind case, realisation) = one `ssrs_step_tracks` launch per GPU
    track_start_region: Tuple[float, float, float, float] = (5, 55, 1, 2)
    track_start_type: str = 'random'       # 'structured' | 'random'
    track_stochastic_nu: float = 1.        # exponent on the move probabilities; != 1 takes the general step, not the fast lane
    track_dirn_restrict: int = 1           # direction memory (0 = whole history, <= 16); 1 is the fast lane's case

    # -- turbines / plotting
    turbine_minimum_hubheight: float = 50.
    turbine_mrkr_styles = ('1k', '2k', '3k', '4k',
                           '+k', 'xk', '*k', '.k', 'ok')
    turbine_mrkr_size: float = 3.
    fig_height: float = 6.
    fig_dpi: int = 200

    def __str__(self):
        # The reference groups instance attributes by *position* in __dict__ (config.py:69-91); section
        # breaks therefore sit at positions 0, 6, 10, 12, 13, 17, 21, 23, 30 of the attribute list.
        names = [f.name for f in fields(Config)]
        breaks = {names.index(first): title for title, first in _SECTIONS}
        # the reference's last break is at position 30 (turbine_minimum_hubheight)
        out = self.__doc__ + '\n'
        for i, key in enumerate(self.__dict__):
            if i in breaks:
                out += f'\n:::: {breaks[i]}\n'
            out += f'{key} = {self.__dict__[key]}\n'
        return out
