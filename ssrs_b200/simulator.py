"""`Simulator` — drop-in for the hot path of `/root/reference/ssrs/simulator.py` (class Simulator(Config)).

Same constructor shape (`Simulator(in_config=None, **kwargs)`), attribute surface, method
names, case/ID strings and on-disk artefacts (`<case>_orograph.npy` f32, `<id>_potential.npy` f32,
`<id>_tracks.pkl`, `summary_presence.npy` f32) as the reference; the bodies that call stages 1-4 run on the GPU:

    reference                                           here
    compute_orographic_updraft_uniform   :189-198       one fused stencil kernel (ssrs_updraft)
    compute_orographic_updrafts_using_wtk :200-215      same kernel with per-cell wind rasters
    _get_interpolated_wind_conditions    :765-792       barycentric interpolation kernel (ssrs_interp_wind)
    compute_thermal_updrafts             :217-228       Philox seeds + separable Gaussian blur on the GPU
    load_updrafts                        :230-243       threshold on the GPU (ssrs_threshold)
    get_directional_potential            :259-288       matrix-free AMG/BiCGStab solve (ssrs_potential_solve)
    simulate_tracks                      :332-386       one batched launch (ssrs_step_tracks) instead of mp.Pool
    plot_presence_map (data part)        :508-546       compute_presence_map(): counts fused into the stepper

Out of scope: terrain/WTK/turbine downloads, CRS handling and matplotlib plotting.  The reference
constructor always downloads terrain; here it is injected with keyword-only extensions:
    elevation=   float raster [rows=north, cols=east] matching `gridsize`
    wind_cases=  {case_id: (wspeed, wdirn)} for snapshot/seasonal modes (stands in for the WTK download): rasters
                 [rows, cols], scalars, or — with wind_points= — 1-D values at the scattered sites, which are then
                 interpolated to the terrain grid on the GPU like the reference's griddata(linear) recipe
    wind_points= (xlocs, ylocs) projected site coordinates (the reference's get_wtk_locs())
    bounds=      projected bounds (west, south, east, north); default puts the south-west corner at (0, 0)
    force_potential=True   recompute every potential even if `<id>_potential.npy` exists.  Independently of it a cached
                 potential is only reused when the content key stored beside it (`<id>_potential.key`: a hash of the
                 thresholded updraft the solve would consume, the movement direction and the grid) matches — the
                 reference's cache (`simulator.py:264-272`) is keyed by the id string alone and silently reuses a
                 potential computed for other terrain or wind; a file without a key counts as stale.
When `torch.distributed` is initialised, tracks are block-partitioned by global id over the ranks, fields are
replicated, the potential solve is row-sharded and presence maps are summed with one all-reduce (NCCL on GPUs).
    case_parallel=True   instead distributes the wind CASES over the ranks (seasonal mode: rank r takes cases
                 r, r + world, ...): every rank runs its cases exactly like a single-GPU run — no solver or presence
                 communication — and only the summary presence map is reduced (the alternative to the
                 row-sharded solve for seasonal throughput).
"""
from __future__ import annotations

import json
import os
import pickle
import time
from dataclasses import asdict
from typing import Dict, Optional

import numpy as np

from . import _native as N
from . import dist as _dist
from . import trackio
from .config import Config
from .layers import get_above_threshold_speed, updraft_fields
from .movmodel import (MovModel, get_starting_indices, interleave_fields, record_tracks_packed,
                       simulate_tracks_batch)
from .potential import solve_potential_device

TRACKS_PKL_LIMIT = 50_000      # above this the pickled list-of-arrays format is impractical; a packed .npz is written


def _elapsed(t0: float) -> str:
    s = time.time() - t0
    return f"{int(s // 60)} min {int(s % 60)} sec" if s >= 60 else f"{s:.2f} sec"


class _ArtefactWriter:
    """Background writer for the on-disk artefacts (.npy/.npz/.pkl): a 120 MB raster takes longer to write than the GPU
    takes to produce the next one, and numpy releases the GIL while writing.  `wait()` before anything reads them."""

    def __init__(self):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=2)
        self._pending = []

    def submit(self, fn, *args, **kwargs):
        self._pending.append(self._pool.submit(fn, *args, **kwargs))

    def wait(self):
        pending, self._pending = self._pending, []
        for f in pending:
            f.result()              # re-raises a failed write


class _SoloDist:
    """`ssrs_b200.dist` as seen by one rank that owns whole cases (case_parallel mode): no sharding, no collectives."""

    @staticmethod
    def rank(): return 0

    @staticmethod
    def world_size(): return 1

    @staticmethod
    def barrier(): return None

    @staticmethod
    def agree(flag): return bool(flag)

    shard_range = staticmethod(_dist.shard_range)

    @staticmethod
    def allreduce_sum(t): return t

    @staticmethod
    def presence_allreduce(t): return t

    @staticmethod
    def gather_tracks(tracks): return tracks


class Simulator(Config):
    """ Class for SSRS simulation """

    lonlat_crs = 'EPSG:4326'
    time_format = 'y%Ym%md%dh%H'

    def __init__(self, in_config: Config = None, *, elevation=None, wind_cases: Optional[Dict] = None,
                 wind_points=None, bounds=None, case_parallel: bool = False, force_potential: bool = False,
                 **kwargs) -> None:
        if in_config is None:
            super().__init__(**kwargs)
        else:
            super().__init__(**asdict(in_config))
        N.load()                 # no CUDA library -> fail here, loudly
        print(f'\n---- SSRS in {self.sim_mode} mode')
        print(f'Run name: {self.run_name}')
        if self.sim_seed >= 0:                                   # reference :50-52
            print('Specified random number seed:', self.sim_seed)
            np.random.seed(self.sim_seed)

        print(f'Output dir: {os.path.join(self.out_dir, self.run_name)}')
        self.data_dir = os.path.join(self.out_dir, self.run_name, 'data/')
        self.fig_dir = os.path.join(self.out_dir, self.run_name, 'figs/')
        self.mode_data_dir = os.path.join(self.data_dir, self.sim_mode)
        self.mode_fig_dir = os.path.join(self.fig_dir, self.sim_mode)
        for d in (self.mode_data_dir, self.mode_fig_dir):
            os.makedirs(d, exist_ok=True)
        # same JSON dump of the configuration as the reference (:63-67); arrays are attached afterwards
        with open(os.path.join(self.out_dir, self.run_name, f'{self.run_name}.json'), 'w', encoding='utf-8') as f:
            json.dump(self.__dict__, f, ensure_ascii=False, indent=2)

        # per-case work sees `self._d`: the process group, or a single-process stand-in when cases are distributed
        self._case_parallel = bool(case_parallel) and _dist.world_size() > 1
        self._d = _SoloDist() if self._case_parallel else _dist
        print(f'Terrain resolution = {self.resolution} m')
        xsize = int(round((self.region_width_km[0] * 1000. / self.resolution)))
        ysize = int(round((self.region_width_km[1] * 1000. / self.resolution)))
        self.gridsize = (ysize, xsize)
        print(f'Terrain grid size = {self.gridsize}')
        if bounds is None:
            bounds = (0.0, 0.0, (xsize - 1) * self.resolution, (ysize - 1) * self.resolution)
        self.bounds = tuple(bounds)
        self.extent = (self.bounds[0], self.bounds[2], self.bounds[1], self.bounds[3])
        self.lonlat_bounds = None            # needs PROJ; not on the hot path
        self.terrain_layers = {'Elevation': 'injected'}
        self.turbines = None
        self.region = None
        # reference :107-127 — attribute surface only: the WTK download is outside the hot path and wind
        # arrives through wind_cases=; case ids that are time stamps give back the reference's `dtimes`
        self.wtk_layers = {
            'wspeed': f'windspeed_{str(int(self.wtk_orographic_height))}m',
            'wdirn': f'winddirection_{str(int(self.wtk_orographic_height))}m',
            'pressure': f'pressure_{str(int(self.wtk_thermal_height))}m',
            'temperature': f'temperature_{str(int(self.wtk_thermal_height))}m',
            'blheight': 'boundary_layer_height',
            'surfheatflux': 'surface_heat_flux',
        }
        self.wtk = None
        self.dtimes = None
        if wind_cases:
            from datetime import datetime
            try:
                self.dtimes = [datetime.strptime(str(k), self.time_format) for k in wind_cases]
            except ValueError:
                self.dtimes = None
        if elevation is None:
            raise ValueError("ssrs_b200.Simulator needs elevation= (terrain download is outside the hot path; "
                             "inject the DEM raster [rows=north, cols=east])")
        torch = N.require_cuda()
        if isinstance(elevation, torch.Tensor):
            z = elevation.to(device="cuda", dtype=torch.float32).contiguous()
        else:
            z = torch.from_numpy(np.ascontiguousarray(elevation, dtype=np.float32)).to("cuda")
        if tuple(z.shape) != self.gridsize:
            raise ValueError(f"elevation shape {tuple(z.shape)} does not match gridsize {self.gridsize}")
        self._elev = z
        if self.movement_model == 'fluidflow':
            # one-time costs of stage 2 paid here instead of inside the first solve: the solver's workspace arena and, in a
            # multi-process run whose solves are row-sharded, the library's NCCL communicator (5 s at 8 GPUs)
            N.check(N.load().ssrs_reserve_workspace(ysize, xsize), "ssrs_reserve_workspace")
            if _dist.world_size() > 1 and not self._case_parallel:
                _dist.native_comm()
        self._writer = _ArtefactWriter()
        self._force_potential = bool(force_potential)
        self._oro_cache: Dict[str, "torch.Tensor"] = {}       # float32 orographs kept on the device (what the .npy holds)
        self._presence: Dict[str, "torch.Tensor"] = {}
        self._track_results = {}
        self.timings: Dict[str, float] = {}

        mode = self.sim_mode.lower()
        if mode != 'uniform':
            if not wind_cases:
                raise ValueError(f"sim_mode={self.sim_mode!r} needs wind_cases= (WTK download is outside the hot path)")
            self.case_ids = list(wind_cases.keys())
            self._wind_cases = wind_cases
            self._wind_points = None
            if wind_points is not None:
                if self.wtk_interp_type not in ('linear', 'nearest', 'cubic'):
                    # griddata's own methods (config.py:60); failing in the constructor beats failing after the terrain work
                    raise ValueError(f"Unknown interpolation method {self.wtk_interp_type!r} for 2 dimensional data")
                from .layers import delaunay_topology, delaunay_triangles
                xl, yl = (np.asarray(v, dtype=np.float64) for v in wind_points)
                topo = delaunay_topology if self.wtk_interp_type == 'cubic' else delaunay_triangles
                self._wind_points = (xl, yl, topo(xl, yl))                     # one triangulation for all cases
            self.compute_orographic_updrafts_using_wtk()
        else:
            print(f'Uniform mode: Wind speed = {self.uniform_windspeed} m/s')
            print(f'Uniform mode: Wind dirn = {self.uniform_winddirn} deg(cw)')
            self.case_ids = [self._get_uniform_id()]
            self.compute_orographic_updraft_uniform()
        for case_id in self._my_case_ids():
            self.compute_thermal_updrafts(case_id)
        fig_aspect = self.region_width_km[0] / self.region_width_km[1]
        self.fig_size = (self.fig_height * fig_aspect, self.fig_height)
        self.km_bar = min([1, 5, 10], key=lambda x: abs(x - self.region_width_km[0] // 4))
        self.flush()             # the constructor's artefacts (orographs, thermals) are on disk when it returns
        print('SSRS Simulator initiation done.')

    def _my_case_ids(self):
        """Cases this rank computes: all of them, or every world-th one in case_parallel mode."""
        if not self._case_parallel:
            return list(self.case_ids)
        return list(self.case_ids[_dist.rank()::_dist.world_size()])

    # ---------------------------------------------------------------- terrain
    def get_terrain_elevation(self):
        return self._elev.cpu().numpy()

    def get_terrain_slope(self):
        return updraft_fields(self._elev, self.resolution, 0.0, 0.0, want=("slope",))["slope"].cpu().numpy()

    def get_terrain_aspect(self):
        return updraft_fields(self._elev, self.resolution, 0.0, 0.0, want=("aspect",))["aspect"].cpu().numpy()

    def get_terrain_layer(self, lname: str):
        return {'Elevation': self.get_terrain_elevation, 'Slope': self.get_terrain_slope,
                'Aspect': self.get_terrain_aspect}[lname]()

    def get_terrain_grid(self):
        xgrid = np.linspace(self.bounds[0], self.bounds[0] + (self.gridsize[1] - 1) * self.resolution, self.gridsize[1])
        ygrid = np.linspace(self.bounds[1], self.bounds[1] + (self.gridsize[0] - 1) * self.resolution, self.gridsize[0])
        return xgrid, ygrid

    # ---------------------------------------------------------------- stage 1
    ORO_CACHE_BYTES = 32 << 30          # device budget for cached orographs (64 cases x 480 MB at 12000 x 10000 = 31 GB)

    def _save_orograph(self, case_id, orograph):
        """`<case>_orograph.npy` (float32, reference :197-198), written in the background; the device copy is kept
        so that load_updrafts' float32 round trip through the file (reference :232-233) costs nothing."""
        if self._d.rank() == 0:
            self._writer.submit(np.save, f'{self._get_orograph_fname(case_id, self.mode_data_dir)}.npy', orograph.cpu().numpy())
        if orograph.numel() * 4 * len(self.case_ids) <= self.ORO_CACHE_BYTES:
            self._oro_cache[case_id] = orograph
        else:
            self._writer.wait()
            self._d.barrier()

    def flush(self):
        """Blocks until every artefact of this Simulator is on disk (collective when torch.distributed is initialised)."""
        self._writer.wait()
        self._d.barrier()

    def compute_orographic_updraft_uniform(self) -> None:
        print('Computing orographic updrafts..')
        t0 = time.time()
        out = updraft_fields(self._elev, self.resolution, float(self.uniform_windspeed), float(self.uniform_winddirn),
                             self.updraft_threshold, want=("orograph",))
        self.timings['updraft_s'] = time.time() - t0
        self._save_orograph(self.case_ids[0], out["orograph"])

    def compute_orographic_updrafts_using_wtk(self) -> None:
        print('Computing orographic updrafts..', end="")
        t0 = time.time()
        for case_id in self._my_case_ids():
            ws, wd = self._wind_cases[case_id]
            if self._wind_points is not None and np.ndim(ws) == 1:
                ws, wd = self._get_interpolated_wind_conditions(ws, wd)
            out = updraft_fields(self._elev, self.resolution, ws, wd, self.updraft_threshold, want=("orograph",))
            self._save_orograph(case_id, out["orograph"])
        print(f'took {_elapsed(t0)}', flush=True)

    def _get_interpolated_wind_conditions(self, wspeed, wdirn):
        """Reference :778-792 — site values -> CUDA rasters (speed, direction in degrees) on the terrain grid."""
        from .layers import interpolate_wind_to_grid
        xl, yl, tri = self._wind_points
        return interpolate_wind_to_grid(xl, yl, wspeed, wdirn, self.bounds[0], self.bounds[1], self.resolution,
                                        self.gridsize, triangles=tri, method=self.wtk_interp_type)

    def get_wtk_locs(self):
        """Reference :711-716 — projected coordinates of the wind sites (here: the injected wind_points=)."""
        if getattr(self, '_wind_points', None) is None:
            raise ValueError("no wind sites: pass wind_points=(xlocs, ylocs) to the constructor")
        return self._wind_points[0], self._wind_points[1]

    def compute_thermal_updrafts(self, case_id: str):
        """Reference :217-228.  Realisation r is keyed by (sim_seed, case, r): reproducible for sim_seed >= 0."""
        if self.thermals_realization_count > 0:
            from .layers import compute_thermals
            print('Computing thermal updrafts...', flush=True)
            aspect = updraft_fields(self._elev, self.resolution, 0.0, 0.0, want=("aspect",))["aspect"]
            ci = self.case_ids.index(case_id)
            for real_id in range(self.thermals_realization_count):
                seed = None if self.sim_seed < 0 else (self.sim_seed * 7919 + ci * 104729 + real_id + 1)
                thermals = compute_thermals(aspect, 2.0, seed=seed)
                if self._d.rank() == 0:
                    np.save(f'{self._get_thermal_fname(case_id, real_id, self.mode_data_dir)}.npy', thermals.cpu().numpy())
            self._d.barrier()
        else:
            print('No thermals requested!', flush=True)

    def load_updrafts(self, case_id: str, apply_threshold=True):
        """[orograph, orograph + thermals_0, ...] (thresholded), float32 numpy (reference :230-243)."""
        return [u.cpu().numpy() for u in self._load_updrafts_device(case_id, apply_threshold)]

    def _load_updrafts_device(self, case_id, apply_threshold=True):
        torch = N.require_cuda()
        oro = self._oro_cache.get(case_id)
        if oro is None:
            oro = torch.from_numpy(np.load(f'{self._get_orograph_fname(case_id, self.mode_data_dir)}.npy')).to("cuda")
        updrafts = [oro]
        for real_id in range(int(self.thermals_realization_count)):
            th = np.load(f'{self._get_thermal_fname(case_id, real_id, self.mode_data_dir)}.npy')
            updrafts.append(oro + torch.from_numpy(th).to("cuda"))
        if apply_threshold:
            updrafts = [get_above_threshold_speed(u, self.updraft_threshold) for u in updrafts]
        return updrafts

    def _get_orograph_fname(self, case_id: str, dirname: str = './'):
        return os.path.join(dirname, f'{case_id}_orograph')

    def _get_thermal_fname(self, case_id: str, real_id: int, dirname: str = './'):
        return os.path.join(dirname, f'{case_id}_r{real_id}_thermals')

    # ---------------------------------------------------------------- stage 2
    def _potential_key(self, updraft) -> str:
        """Content key of a potential: what the solve consumes (thresholded updraft raster, direction, grid)."""
        import hashlib
        torch = N.require_cuda()
        u = updraft if isinstance(updraft, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(updraft, dtype=np.float32))
        u = u.to(device="cuda", dtype=torch.float32).contiguous()
        # two independent 64-bit sums of the raster's bit patterns, computed on the device (a 120 MB raster would take
        # longer to hash on the host than to solve)
        bits = u.view(torch.int32).to(torch.int64).ravel()
        idx = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64)
        s1 = int(bits.sum().item()) & (2 ** 64 - 1)
        s2 = int((bits * (idx % 1000003 + 1)).sum().item()) & (2 ** 64 - 1)
        raw = f"{self.gridsize}|{float(self.track_direction)!r}|{s1:x}|{s2:x}"
        return hashlib.sha256(raw.encode()).hexdigest()[:32]

    def get_directional_potential(self, updraft, case_id, real_id):
        """float32 potential for (case, realisation); cached as `<id>_potential.npy` like the reference
        (:259-288, including its cache rules) plus a content key (see the class docstring).  `updraft` may be numpy or
        a CUDA tensor.  An unconverged solve raises and is never cached."""
        torch = N.require_cuda()
        fname = self._get_potential_fname(case_id, real_id, self.mode_data_dir)
        id_str = self._get_id_string(case_id, real_id)
        key = self._potential_key(updraft)
        try:
            usable = (not self._force_potential) and os.path.exists(f'{fname}.npy') and os.path.exists(f'{fname}.key')
            if usable:
                with open(f'{fname}.key') as f:
                    usable = f.read().strip() == key
            if not self._d.agree(usable):
                raise FileNotFoundError         # the solve is collective: all ranks follow rank 0's view of the cache
            potential = np.load(f'{fname}.npy')
            if potential.shape != self.gridsize:
                raise FileNotFoundError
            if (self.sim_seed < 0) & (real_id != 0):
                raise FileNotFoundError
            print(f'{id_str}: Found saved potential')
            pot_dev = torch.from_numpy(potential).to("cuda")
        except FileNotFoundError:
            t0 = time.time()
            print(f'{id_str}: Computing potential..', end="", flush=True)
            pot_dev, stats = solve_potential_device(updraft, self.track_direction, strict=False,
                                                    sharded=self._d.world_size() > 1)       # row-sharded over the ranks
            self.timings['potential_s'] = time.time() - t0
            self.solve_stats = stats
            print(f'took {_elapsed(t0)}', flush=True)
            if not stats['converged']:
                # the reference's direct solve cannot fail this way; a wrong field must neither be stepped on nor cached
                raise N.NativeError(f"{id_str}: potential solve did not converge (relative residual "
                                    f"{stats['rel_residual']:.3e} after {stats['iterations']} iterations); nothing was cached")
            potential = pot_dev.cpu().numpy()
            if self._d.rank() == 0:
                self._writer.submit(self._save_potential, fname, potential, key)
        if np.isnan(potential).any():
            print('NANs found in potential!')
        self._last_potential_device = pot_dev
        return potential

    @staticmethod
    def _save_potential(fname, potential, key):
        np.save(f'{fname}.npy', potential)
        with open(f'{fname}.key', 'w') as f:       # written after the raster: a key without a complete raster cannot exist
            f.write(key + '\n')

    def _get_id_string(self, case_id: str, real_id: Optional[int] = None):
        out = f'{case_id}_d{int(self.track_direction % 360)}_t{int(self.updraft_threshold * 100)}_{self.movement_model}'
        if real_id is not None:
            out += f'_r{int(real_id)}'
        return out

    def _get_potential_fname(self, case_id: str, real_id: int, dirname: str):
        return os.path.join(dirname, f'{self._get_id_string(case_id, real_id)}_potential')

    # ---------------------------------------------------------------- stage 3 + 4
    def _track_seed(self, case_index: int, real_id: int) -> int:
        base = self.sim_seed if self.sim_seed >= 0 else int.from_bytes(os.urandom(4), 'little')
        return (base * 1_000_003 + case_index * 1009 + real_id) & (2 ** 63 - 1)

    STEPS_IN_FLIGHT = 4                 # (case, realisation) stepping launches that may overlap on their own streams

    def simulate_tracks(self, save_tracks: Optional[bool] = None):
        """Simulate tracks (reference :332-386).  Presence counts are accumulated on the device during
        stepping and kept per (case, realisation) in `self.presence_counts(id)`.

        The reference steps one (case, realisation) after the other and waits for each pool.  Here every stepping launch
        goes to one of a few side streams (phased launch, `ssrs_step_tracks_phased`) and its presence all-reduce to a
        reduce stream, so the long tail of one launch — it lasts as long as its longest track — overlaps the next
        case's potential solve and the bulk of its stepping; trajectories (when recorded) and files are produced after
        the launches have drained."""
        torch = N.require_cuda()
        print(f'Movement model = {self.movement_model}')
        print(f'Updraft threshold = {self.updraft_threshold} m/s')
        print(f'Movement direction = {self.track_direction} deg (cw)')
        starting_rows, starting_cols = get_starting_indices(
            self.track_count, self.track_start_region, self.track_start_type, self.region_width_km, self.resolution)
        n = len(starting_rows)
        lo, hi = self._d.shard_range(n, self._d.rank(), self._d.world_size())
        record = (n <= TRACKS_PKL_LIMIT) if save_tracks is None else bool(save_tracks)
        if self.movement_model not in ('fluidflow', 'drw'):
            raise ValueError(f'Invalid movement_model {self.movement_model!r}; options: fluidflow, drw')
        # several launches in flight: phased launches (survivors compacted, so a launch's tail leaves the SMs to the next
        # one); a single launch is fastest stepped to the end in one go (its tracks run alone, one per lane)
        n_launches = len(self._my_case_ids()) * (1 + int(self.thermals_realization_count))
        phased = self.track_dirn_restrict == 1 and n_launches > 1
        slots = self.STEPS_IN_FLIGHT
        streams = [torch.cuda.Stream() for _ in range(slots)]
        reduce_stream = torch.cuda.Stream() if self._d.world_size() > 1 else None
        in_slot = [None] * slots            # what the slot's last launch still reads (fields, workspace, result)
        launched = []
        t_all = time.time()
        for case_id in self._my_case_ids():
            ci = self.case_ids.index(case_id)
            for real_id, updraft in enumerate(self._load_updrafts_device(case_id)):
                if self.sim_seed > 0:
                    np.random.seed(self.sim_seed + real_id)              # reference :351-352
                id_str = self._get_id_string(case_id, real_id)
                fields = None
                if self.movement_model == 'fluidflow':
                    self.get_directional_potential(updraft, case_id, real_id)
                    fields = interleave_fields(updraft, self._last_potential_device)
                print(f'{id_str}: Simulating {self.track_count} tracks..', flush=True)
                k = len(launched) % slots
                s = streams[k]
                if in_slot[k] is not None:
                    s.synchronize()                                   # the slot's previous launch has drained: its buffers go
                    in_slot[k] = None
                s.wait_stream(torch.cuda.current_stream())            # fields are ready
                with torch.cuda.stream(s):
                    res = simulate_tracks_batch(self.track_direction, starting_rows[lo:hi], starting_cols[lo:hi],
                                                self.gridsize, self.track_dirn_restrict, self.track_stochastic_nu,
                                                fields=fields, seed=self._track_seed(ci, real_id), track_id0=lo,
                                                phased=phased)
                    steps = res._total
                    if reduce_stream is None:
                        presence = res.presence
                if reduce_stream is not None:
                    reduce_stream.wait_stream(s)
                    with torch.cuda.stream(reduce_stream):            # every rank issues the collectives in the same order
                        presence = self._d.presence_allreduce(res.presence)      # ssrs_presence_allreduce (NCCL)
                        steps = self._d.allreduce_sum(res._total.clone())
                in_slot[k] = (fields, res)
                launched.append((case_id, real_id, ci, id_str, fields if record else None, res, presence, steps))
        for s in streams:
            s.synchronize()
        if reduce_stream is not None:
            reduce_stream.synchronize()
        self.timings['tracks_s'] = time.time() - t_all
        print(f'Simulating tracks took {_elapsed(t_all)}', flush=True)
        for case_id, real_id, ci, id_str, fields, res, presence, steps in launched:
            self.total_track_steps = int(steps.item())
            self._presence[id_str] = presence
            fname = self._get_tracks_fname(case_id, real_id, self.mode_data_dir)
            if record:
                # trajectories: a second, recording pass in chunks sized by the now known lengths (the longest track
                # is ~10x the mean, so one dense [longest, n] buffer would be mostly padding); the counter-based
                # streams make it repeat the counting pass step for step
                off, pts = record_tracks_packed(self.track_direction, starting_rows[lo:hi], starting_cols[lo:hi],
                                                self.gridsize, self.track_dirn_restrict, self.track_stochastic_nu,
                                                fields=fields, seed=self._track_seed(ci, real_id), track_id0=lo,
                                                lengths=res.traj_len.cpu().numpy())
                if n <= TRACKS_PKL_LIMIT:
                    tracks = self._d.gather_tracks([a.copy() for a in trackio.unpack_tracks(off, pts)])
                    if self._d.rank() == 0:
                        self._writer.submit(trackio.save_tracks_pickle, fname, tracks)    # the reference's file (:383-386)
                else:
                    # large runs: packed offsets + points (trackio.py); one file per rank's block of track ids
                    suffix = '' if self._d.world_size() == 1 else f'_part{self._d.rank()}of{self._d.world_size()}'
                    self._writer.submit(trackio.save_tracks_packed, f'{fname}{suffix}', off, pts)
            if self._d.rank() == 0:
                # counts of every run are kept on disk (uncompressed: compressing a 120 MB raster costs seconds), so
                # that a fresh Simulator over this run directory can build the presence map without the tracks
                self._writer.submit(np.savez, f'{fname}_presence_counts.npz', counts=presence.cpu().numpy())
        self.flush()

    def load_tracks(self, case_id: Optional[str] = None, real_id: int = 0):
        """The stored tracks of (case, realisation) as the reference's list of int16 [L, 2] arrays, from either
        on-disk format (single-rank runs)."""
        case_id = self.case_ids[0] if case_id is None else case_id
        return trackio.load_tracks(self._get_tracks_fname(case_id, real_id, self.mode_data_dir))

    def _presence_device(self, case_id: str, real_id: int):
        """Counts of (case, realisation) on the device: from this object's last simulate_tracks(), else from the run
        directory (`<id>_tracks_presence_counts.npz`, or recounted from the stored tracks like the reference, which
        re-reads `<id>_tracks.pkl`, simulator.py:525-529)."""
        torch = N.require_cuda()
        key = self._get_id_string(case_id, real_id)
        if key not in self._presence:
            fname = self._get_tracks_fname(case_id, real_id, self.mode_data_dir)
            if os.path.exists(f'{fname}_presence_counts.npz'):
                with np.load(f'{fname}_presence_counts.npz') as z:
                    counts = z['counts']
            else:
                try:
                    tracks = trackio.load_tracks(fname)
                except FileNotFoundError:
                    raise KeyError(f"no presence counts for {key}: not simulated by this Simulator and neither "
                                   f"{fname}_presence_counts.npz nor stored tracks exist in the run directory") from None
                from .movmodel import compute_presence_counts
                counts = compute_presence_counts(tracks, self.gridsize)
            self._presence[key] = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int32)).to("cuda")
        return self._presence[key]

    def presence_counts(self, case_id: Optional[str] = None, real_id: int = 0) -> np.ndarray:
        """int32 visit counts of (case, realisation) (reference compute_presence_counts, movmodel.py:410-419)."""
        case_id = self.case_ids[0] if case_id is None else case_id
        return self._presence_device(case_id, real_id).cpu().numpy()

    def _get_tracks_fname(self, case_id: str, real_id: int, dirname: str):
        return os.path.join(dirname, f'{self._get_id_string(case_id, real_id)}_tracks')

    def _get_presence_fname(self, case_id: str, real_id: int, dirname: str):
        return os.path.join(dirname, f'{self._get_id_string(case_id, real_id)}_presence')

    def compute_presence_map(self, radius: float = 1000.) -> np.ndarray:
        """Data part of the reference's plot_presence_map (:508-546): smooth each realisation's counts with the
        disk kernel, normalise by the maxima, sum over cases, save `summary_presence.npy` (float32)."""
        from .presence import smooth_presence_counts
        krad = min(max(radius / self.resolution, 2), min(self.gridsize) / 2)
        summary = None
        for case_id in self._my_case_ids():
            case_prob = None
            for real_id in range(1 + int(self.thermals_realization_count)):
                counts = self._presence_device(case_id, real_id)
                pr = smooth_presence_counts(counts, int(round(krad)))
                pr = pr / pr.max()
                case_prob = pr if case_prob is None else case_prob + pr
            case_prob = case_prob / case_prob.max()
            summary = case_prob if summary is None else summary + case_prob
        if self._case_parallel:                 # the only exchange of this mode: sum of the ranks' case maps
            torch = N.require_cuda()
            if summary is None:                 # more ranks than cases
                summary = torch.zeros(self.gridsize, dtype=torch.float32, device="cuda")
            summary = _dist.allreduce_sum(summary.float().contiguous())
        summary = (summary / summary.max()).float().cpu().numpy()
        if _dist.rank() == 0:
            np.save(os.path.join(self.mode_data_dir, 'summary_presence.npy'), summary)
        return summary

    def plot_presence_map(self, plot_turbs=True, radius: float = 1000., show=False, minval=0.1, plot_all: bool = False):
        """Computes and saves the summary presence raster; figures are outside the hot path (no matplotlib)."""
        print('Computing presence density map (figures are not produced by ssrs_b200)..')
        return self.compute_presence_map(radius)

    def _no_plots(self, *a, **k):
        raise NotImplementedError("plotting is outside the B200 hot path (SURVEY.md §2 row 5); the data products "
                                  "(.npy/.pkl) are written with the reference's names, so the reference's plot_* "
                                  "methods can read them")

    plot_directional_potentials = plot_simulated_tracks = plot_updrafts = plot_wtk_layers = _no_plots
    plot_terrain_features = plot_terrain_elevation = plot_terrain_slope = plot_terrain_aspect = _no_plots
    plot_windplant_presence_map = plot_updraft_threshold_function = _no_plots

    def _get_uniform_id(self):
        return f's{int(self.uniform_windspeed)}d{int(self.uniform_winddirn)}'
