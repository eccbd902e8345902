"""Host-side mirror of ``uclchem.model`` for the GPU path.

Same names, argument meaning and return conventions as the reference
(``src/uclchem/model.py``: ``cloud`` :227-316, ``collapse`` :319-426, ``hot_core`` :429-524,
``cshock`` :527-644), but the work goes through the C ABI of ``include/uclgpu.h`` instead of
``uclchemwrap``.  On top of the per-model functions the module adds what the
reference leaves to user scripts (``scripts/grid.py:41-59``): ``cloud_grid``,
``hot_core_grid`` and ``cshock_grid`` integrate a whole table of models in one call.

Differences from the reference, all deliberate:

* file based I/O is done on the host after the run (``datio.py``): ``outputFile`` (full output,
  ``io.f90`` formats 335/8020), ``abundSaveFile`` / ``abundLoadFile`` (format 8010) are honoured in the
  reference's disk mode (neither ``return_array`` nor ``return_dataframe``), as are ``columnFile`` (+ ``writeStep``)
  and ``rateFile``; ``fluxFile`` is not written; combining file keys with the in-memory modes is
  refused with the reference's own error (``model.py:121-131``);
* parameters start from ``defaultparameters.f90`` on every call (the reference leaks
  parameters between calls);
* a model failure is reported through the success flag, never raised (``utils.check_error`` semantics,
  ``utils.py:30-51``); an unknown parameter key gives flag -1 (PARAMETER_READ_ERROR) for single models, while the
  ``*_grid`` functions raise ``KeyError`` before anything runs (a table of models has no single flag).
"""
from __future__ import annotations

import numpy as np

from . import datio
from ._capi import N_PHYS, get_library
from .params import MODEL_KINDS, params_from_dict

TIMEPOINTS = 500  # src/uclchem/constants.py
PHYSICAL_PARAMETERS = ["Time", "Density", "gasTemp", "dustTemp", "Av", "radfield", "zeta", "point"]


def _lower(param_dict):
    out = {}
    for k, v in (param_dict or {}).items():
        assert k.lower() not in out, f"Lower case key {k} is already in the dict, stopping"
        out[k.lower()] = v
    return out


def pre_flight_checklist(return_array, return_dataframe, return_rates, starting_chemistry=None, user_params=None):
    """model.py:103-155: in-memory and disk modes are not mixed.  (The reference also forbids switching
    mode within one Python session -- an artefact of its SAVEd Fortran file units; not needed here.)"""
    user_params = user_params or {}
    if starting_chemistry is not None:
        assert return_array or return_dataframe, (
            "starting_chemistry can only be used with return_array or return_dataframe set to True;\n"
            "Instead specify 'abundLoadFile' in the param_dict to load starting abundances from a file.")
    if return_array or return_dataframe or return_rates:
        file_keys = [k for k in user_params if k.lower().endswith("file")]
        if file_keys:
            raise RuntimeError("return_array or return_dataframe cannot be used if any output of input file is "
                               "specified.\n" + f"Offending keys: {', '.join(file_keys)}")
        if return_rates:
            assert return_array or return_dataframe, (
                "return_rates and return_heating can only be used with return_array or return_dataframe set to True; ")


def _coefficients(pd_):
    """The reference's per-reaction overrides (wrap.f90:744-761): `alpha`, `beta`, `gamma` entries of the parameter
    dictionary are {reaction index (1-based, as in network.f90): value} dictionaries.  Returned as the C ABI's
    (which, 0-based reaction, value) triples; the keys are removed from `pd_`."""
    out = []
    for which, key in enumerate(("alpha", "beta", "gamma")):
        d = pd_.pop(key, None)
        if d is None:
            continue
        if not isinstance(d, dict):
            raise ValueError(f"{key} must be a dictionary of reaction index: value pairs")
        out += [(which, int(k) - 1, float(v)) for k, v in d.items()]
    return out


def _format_output(n_out, abunds, success_flag):
    """model.py:76-81"""
    abunds = [] if (success_flag < 0 or n_out == 0) else list(abunds[:n_out])
    return [int(success_flag)] + abunds


def _run_single(kind, param_dict, out_species, return_array, return_dataframe, return_rates, starting_chemistry,
                timepoints, extra, **run_kw):
    lib = get_library()
    pd_ = _lower(param_dict)
    traj = return_array or return_dataframe
    pre_flight_checklist(return_array, return_dataframe, return_rates, starting_chemistry, pd_)
    # disk mode of the reference: files named in the dictionary are read before / written after the run
    files = {k: pd_.pop(k) for k in list(pd_) if k.endswith("file")}
    if "fluxfile" in files:   # REACTIONRATE is only filled by networks built with MakeRates' enable_rates_to_disk
        raise NotImplementedError("fluxFile is not written by the GPU path; use return_rates and "
                                  "uclchem_b200.analysis.rates_to_dy_and_flux")
    if "columnfile" in files and not out_species:
        raise ValueError("columnFile needs out_species (the reference writes the species of outSpecies)")
    write_step = int(pd_.pop("writestep", 1))
    pd_.update(extra)
    coefficients = _coefficients(pd_)
    try:
        params = params_from_dict(pd_, ncell=1)
    except KeyError as e:
        # wrap.f90:966-970: an unknown key makes dictionaryParser return PARAMETER_READ_ERROR; the model is not
        # run and the caller sees the flag (utils.check_error), not an exception
        print(f"Parameter read failed: {e.args[0]}")
        return _failed_result(kind, traj, -1)
    y0 = None
    if starting_chemistry is not None:
        sc = np.asarray(starting_chemistry, dtype=np.float64).ravel()
        y0 = np.zeros((1, lib.neq))
        y0[0, : lib.nspec] = sc[: lib.nspec]
    elif "abundloadfile" in files:          # readInputAbunds, io.f90:36-46
        y0 = np.zeros((1, lib.neq))
        y0[0, : lib.nspec] = datio.read_abundances(files["abundloadfile"], lib.nspec)
    want_rows = traj or any(k in files for k in ("outputfile", "columnfile", "ratefile"))
    want_rates = (traj and return_rates) or "ratefile" in files
    out = lib.run_grid(MODEL_KINDS[kind], params, y0=y0, timepoints=timepoints if want_rows else 0,
                       want_physics=want_rows, want_chem=want_rows, want_rates=want_rates, coefficients=coefficients,
                       **run_kw)
    while not traj and want_rows and int(out["flag"][0]) == -6 and timepoints < (1 << 20):
        # disk mode has no row limit in the reference (rows go straight to the file, io.f90:59-83): the rows come
        # back through the in-memory buffers here, so a model with more output intervals is re-run with more room
        timepoints *= 4
        out = lib.run_grid(MODEL_KINDS[kind], params, y0=y0, timepoints=timepoints, want_physics=True, want_chem=True,
                           want_rates=want_rates, coefficients=coefficients, **run_kw)
    flag = int(out["flag"][0])
    tdiss = float(out["dissipation_time"][0]) if flag >= 0 else None   # model.py:606-607
    if not traj:
        nrows = min(int(out["stats"][0][7]) + 1, timepoints + 1)
        if "outputfile" in files:
            datio.write_full_output(files["outputfile"], lib.species, out["physics"][0, :nrows],
                                    out["abund"][0, :nrows])
        if "columnfile" in files:
            datio.write_column_output(files["columnfile"], lib.species, out_species, out["physics"][0, :nrows],
                                      out["abund"][0, :nrows], write_step)
        if "ratefile" in files:
            datio.write_rate_output(files["ratefile"], out["physics"][0, :nrows], out["rates"][0, :nrows])
        if "abundsavefile" in files:          # finalOutput, io.f90:48-56
            datio.write_abundances(files["abundsavefile"], out["y_final"][0, : lib.nspec])
        n_out = len(out_species) if out_species else 0
        idx = [lib.species.index(s) for s in (out_species or [])]
        res = _format_output(n_out, out["y_final"][0, idx], flag)
        if kind == "cshock":
            res = [res[0], tdiss] + res[1:]   # model.py:643
        return res
    nrows = int(out["stats"][0][7]) + 1  # row 0 + one row per interval (model.py:158-187 trims by Time != 0)
    nrows = min(nrows, timepoints + 1)
    physics = out["physics"][0, :nrows][:, None, :]
    chem = out["abund"][0, :nrows][:, None, :]
    rates = out["rates"][0, :nrows][:, None, :] if return_rates else None
    abundance_start = chem[nrows - 1, 0, :].copy()
    if return_dataframe:
        import pandas as pd

        physics = pd.DataFrame(physics[:, 0, :N_PHYS], columns=PHYSICAL_PARAMETERS)
        chem = pd.DataFrame(chem[:, 0, :], columns=lib.species)
        if rates is not None:
            rates = pd.DataFrame(rates[:, 0, :])
    if kind == "cshock":   # model.py:618-640: (physics, chem, rates, dissipation_time, abundanceStart, flag)
        return (physics, chem, rates, tdiss, abundance_start, flag)
    return (physics, chem, rates, abundance_start, flag)


def _failed_result(kind, traj, flag):
    """What the reference's wrappers return when the model never ran (empty trajectories, model.py:76-81,606-643)."""
    if not traj:
        return [flag, None] if kind == "cshock" else [flag]
    if kind == "cshock":
        return (None, None, None, None, None, flag)
    return (None, None, None, None, flag)


def cloud(param_dict=None, out_species=None, return_array=False, return_dataframe=False, return_rates=False,
          starting_chemistry=None, timepoints=TIMEPOINTS):
    """Static or free-fall cloud (model.py:227-316)."""
    return _run_single("cloud", param_dict, out_species, return_array, return_dataframe, return_rates,
                       starting_chemistry, timepoints, {})


COLLAPSE_MODES = {"BE1.1": 1, "BE4": 2, "filament": 3, "ambipolar": 4}   # model.py:361


def _collapse_mode(collapse):
    try:
        return COLLAPSE_MODES[collapse]
    except (KeyError, TypeError):
        raise ValueError("collapse must be one of 'BE1.1', 'BE4', 'filament', or 'ambipolar'")


def collapse(collapse, physics_output, param_dict=None, out_species=None, return_array=False, return_dataframe=False,
             return_rates=False, starting_chemistry=None, timepoints=TIMEPOINTS):
    """Collapsing prestellar core, Priestley et al. 2018 (model.py:319-426, collapse.f90).  `collapse` is one of
    'BE1.1', 'BE4', 'filament', 'ambipolar'; the Bonnor-Ebert modes set their own final time (0.97 of the fit's
    time span).  `physics_output` (the reference's per-interval dump of radius / density / velocity to a text
    file) is not written by the GPU path: pass None and read the density from the physics trajectory."""
    mode = _collapse_mode(collapse)
    if physics_output is not None:
        raise NotImplementedError("physics_output is not written by the GPU path; use return_array / return_dataframe "
                                  "and read Density from the physics output")
    return _run_single("collapse", param_dict, out_species, return_array, return_dataframe, return_rates,
                       starting_chemistry, timepoints, {"collapse_mode": mode})


def hot_core(temp_indx, max_temperature, param_dict=None, out_species=None, return_array=False,
             return_dataframe=False, return_rates=False, starting_chemistry=None, timepoints=TIMEPOINTS):
    """Hot core / hot corino warm-up (model.py:429-524)."""
    return _run_single("hot_core", param_dict, out_species, return_array, return_dataframe, return_rates,
                       starting_chemistry, timepoints, {"temp_indx": temp_indx, "max_temperature": max_temperature})


def cshock(shock_vel, timestep_factor=0.01, minimum_temperature=0.0, param_dict=None, out_species=None,
           return_array=False, return_dataframe=False, return_rates=False, starting_chemistry=None,
           timepoints=TIMEPOINTS):
    """C-type shock (model.py:527-644); returns the dissipation time like the reference."""
    return _run_single("cshock", param_dict, out_species, return_array, return_dataframe, return_rates,
                       starting_chemistry, timepoints,
                       {"shock_vel": shock_vel, "timestep_factor": timestep_factor,
                        "minimum_temperature": minimum_temperature})


def jshock(shock_vel, param_dict=None, out_species=None, return_array=False, return_dataframe=False,
           return_rates=False, starting_chemistry=None, timepoints=TIMEPOINTS):
    """J-type shock, James et al. 2020 (model.py:647-745, jshock.f90).  Unlike cshock no dissipation time is
    returned: the tuples are those of `cloud`."""
    return _run_single("jshock", param_dict, out_species, return_array, return_dataframe, return_rates,
                       starting_chemistry, timepoints, {"shock_vel": shock_vel})


def _tracer_history(time_array, density_array, gas_temperature_array, dust_temperature_array, zeta_array,
                    radfield_array, coldens_H_array, coldens_H2_array, coldens_CO_array, coldens_C_array):
    """[..., 10, ntime] history block of the C ABI from the reference's postprocess arguments (model.py:748-830:
    every array must have the length of time_array; times are in seconds like the reference's `timegrid`)."""
    t = np.asarray(time_array, dtype=np.float64)
    cols = [coldens_H_array, coldens_H2_array, coldens_CO_array, coldens_C_array]
    use = coldens_H_array is not None
    if use and any(c is None for c in cols):
        raise ValueError("coldens_H_array, coldens_H2_array, coldens_CO_array and coldens_C_array must be given together")
    rows = [t, density_array, gas_temperature_array, dust_temperature_array, radfield_array, zeta_array]
    rows += cols if use else [np.zeros_like(t)] * 4
    for r in rows:
        assert r is not None and np.shape(r) == t.shape, "All arrays must be the same length"
    return np.stack([np.asarray(r, dtype=np.float64) for r in rows], axis=-2), use


def postprocess(param_dict=None, out_species=None, return_array=False, return_dataframe=False, return_rates=False,
                starting_chemistry=None, time_array=None, density_array=None, gas_temperature_array=None,
                dust_temperature_array=None, zeta_array=None, radfield_array=None, coldens_H_array=None,
                coldens_H2_array=None, coldens_CO_array=None, coldens_C_array=None):
    """Chemistry along a supplied tracer history (model.py:748-880, postprocess.f90): density, temperatures, radiation
    field and cosmic-ray rate -- and optionally the shielding column densities -- are read from the arrays at every
    time of `time_array` (seconds); one output row per history point."""
    grid, use = _tracer_history(time_array, density_array, gas_temperature_array, dust_temperature_array, zeta_array,
                                radfield_array, coldens_H_array, coldens_H2_array, coldens_CO_array, coldens_C_array)
    return _run_single("postprocess", param_dict, out_species, return_array, return_dataframe, return_rates,
                       starting_chemistry, grid.shape[-1], {}, pp_grid=grid[None], pp_coldens=use)


# ---------------------------------------------------------------------------------------
# grids: what scripts/grid.py does with a process pool, in one call
# ---------------------------------------------------------------------------------------
def _run_grid(kind, param_dict, starting_chemistry, extra, out_species=None, return_array=False,
              return_rates=False, timepoints=TIMEPOINTS, **run_kw):
    lib = get_library()
    pd_ = _lower(param_dict)
    file_keys = [k for k in pd_ if k.endswith("file")]
    if file_keys:   # one file name cannot hold a grid; the reference's grid scripts build one name per model
        raise RuntimeError("file output is per model; use return_array=True for grids.\n"
                           f"Offending keys: {', '.join(file_keys)}")
    pd_.update(extra)
    coefficients = _coefficients(pd_)   # one set of overrides for the whole grid
    params = params_from_dict(pd_)
    ncell = params.shape[1]
    y0 = None
    if starting_chemistry is not None:
        sc = np.asarray(starting_chemistry, dtype=np.float64)
        if sc.ndim == 1:
            sc = np.broadcast_to(sc, (ncell, sc.shape[0]))
        y0 = np.zeros((ncell, lib.neq))
        y0[:, : lib.nspec] = sc[:, : lib.nspec]
    out = lib.run_grid(MODEL_KINDS[kind], params, y0=y0, timepoints=timepoints if return_array else 0,
                       want_physics=return_array, want_chem=return_array, want_rates=return_array and return_rates,
                       coefficients=coefficients, **run_kw)
    res = {"flag": out["flag"], "abundances": out["y_final"][:, : lib.nspec], "physics": out["phys_final"],
           "stats": out["stats"], "species": lib.species}
    if return_array:
        # the reference's in-memory layout (wrap.f90:549-560): [time, point, field] with one "point" per cell;
        # rows past a model's last output time stay zero, `nrows` says how many are filled per cell
        res["physics_array"] = np.ascontiguousarray(np.swapaxes(out["physics"], 0, 1))
        res["chemical_abun_array"] = np.ascontiguousarray(np.swapaxes(out["abund"], 0, 1))
        if return_rates:
            res["rates_array"] = np.ascontiguousarray(np.swapaxes(out["rates"], 0, 1))
        res["nrows"] = np.minimum(out["stats"][:, 7] + 1, timepoints + 1)
    if out_species:
        res["out_species"] = out["y_final"][:, [lib.species.index(s) for s in out_species]]
    if kind == "cshock":
        res["dissipation_time"] = out["dissipation_time"]
    return res


def cloud_grid(param_dict, starting_chemistry=None, out_species=None, return_array=False, return_rates=False,
               timepoints=TIMEPOINTS):
    """Integrate a grid of cloud models.  Array-valued entries of ``param_dict`` are per-cell
    columns, scalars broadcast.  Returns a dict with per-cell ``flag`` (constants.f90 codes),
    final ``abundances`` [ncell, nspec], final ``physics`` [ncell, 8] and solver ``stats``; with
    ``return_array`` also the trajectories in the reference's in-memory layout
    (``physics_array`` [timepoints+1, ncell, 8], ``chemical_abun_array`` [timepoints+1, ncell, nspec],
    ``rates_array`` with ``return_rates``) and ``nrows``, the number of filled rows per cell."""
    return _run_grid("cloud", param_dict, starting_chemistry, {}, out_species, return_array, return_rates, timepoints)


def collapse_grid(collapse, param_dict, starting_chemistry=None, out_species=None, return_array=False,
                  return_rates=False, timepoints=TIMEPOINTS):
    """A grid of collapse models; `collapse` is one name or one name per cell."""
    modes = [_collapse_mode(c) for c in collapse] if not isinstance(collapse, str) else _collapse_mode(collapse)
    return _run_grid("collapse", param_dict, starting_chemistry, {"collapse_mode": modes}, out_species, return_array,
                     return_rates, timepoints)


def hot_core_grid(temp_indx, max_temperature, param_dict, starting_chemistry=None, out_species=None,
                  return_array=False, return_rates=False, timepoints=TIMEPOINTS):
    return _run_grid("hot_core", param_dict, starting_chemistry,
                     {"temp_indx": temp_indx, "max_temperature": max_temperature}, out_species, return_array,
                     return_rates, timepoints)


def cshock_grid(shock_vel, param_dict, timestep_factor=0.01, minimum_temperature=0.0, starting_chemistry=None,
                out_species=None, return_array=False, return_rates=False, timepoints=TIMEPOINTS):
    return _run_grid("cshock", param_dict, starting_chemistry,
                     {"shock_vel": shock_vel, "timestep_factor": timestep_factor,
                      "minimum_temperature": minimum_temperature}, out_species, return_array, return_rates,
                     timepoints)


def jshock_grid(shock_vel, param_dict, starting_chemistry=None, out_species=None, return_array=False,
                return_rates=False, timepoints=TIMEPOINTS):
    return _run_grid("jshock", param_dict, starting_chemistry, {"shock_vel": shock_vel}, out_species, return_array,
                     return_rates, timepoints)


def postprocess_grid(param_dict, time_array, density_array, gas_temperature_array, dust_temperature_array, zeta_array,
                     radfield_array, coldens_H_array=None, coldens_H2_array=None, coldens_CO_array=None,
                     coldens_C_array=None, starting_chemistry=None, out_species=None, return_array=False,
                     return_rates=False):
    """A table of tracers in one call: every history array is [ncell, ntime] (all tracers share ntime)."""
    grid, use = _tracer_history(time_array, density_array, gas_temperature_array, dust_temperature_array, zeta_array,
                                radfield_array, coldens_H_array, coldens_H2_array, coldens_CO_array, coldens_C_array)
    assert grid.ndim == 3, "history arrays must be [ncell, ntime]"
    pd_ = dict(param_dict or {})
    pd_.setdefault("initialDens", grid[:, 1, 0])     # one parameter column of the right length fixes ncell
    return _run_grid("postprocess", pd_, starting_chemistry, {}, out_species, return_array, return_rates, grid.shape[-1],
                     pp_grid=grid, pp_coldens=use)
