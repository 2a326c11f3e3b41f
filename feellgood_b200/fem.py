"""Host-side mirror of the reference's time-integration loop around the B200 hot path.

    TimeStepper   src/time_integration.cpp:11-38
    LogStats      src/log-stats.h
    Stats         src/time_integration.cpp:41-47
    Fem           src/fem.h (vmax, E, Etot, Etot0, energy, evolution, compute_all, saver) and
                  Fem::time_integration, src/time_integration.cpp:133-245

Same names, control flow, accept/reject rules and rounding precautions as the reference, so that
a run driven through this loop takes the same sequence of time steps as the reference would.
Everything numerical happens on the GPU through LinAlgebra (C ABI): the per-step hot path, the
energies (Fem::energy), the averages and the maximum angle written to the .evol file.  The demag
solver (ScalFMM in the reference) is outside the path: it is a callback that reads the new
magnetisation and writes phi / phiv, exactly where Fem::compute_all calls myFMM.calc_demag.
"""
from __future__ import annotations

import math
import sys

import numpy as np

FLT_EPSILON = 1.1920928955078125e-07
ENERGY_TERMS = ("EXCHANGE", "ANISOTROPY", "DEMAG", "ZEEMAN")   # src/fem.h:30-36
EXCHANGE, ANISOTROPY, DEMAG, ZEEMAN = range(4)


def round_half_away(x):
    """std::round for the step count: round half away from zero (Python's round() rounds half to even)."""
    import math
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


class TimeStepper:
    """Logic for finding a reasonable time step (src/time_integration.cpp:11-38)."""

    def __init__(self, initial, dtmin, dtmax):
        self.hard_min = dtmin
        self.hard_max = dtmax * (1 + FLT_EPSILON)
        self.soft_max = initial

    def set_soft_limit(self, mx):
        self.soft_max = min(self.soft_max, mx)

    def __call__(self, stride):
        step = min(stride, self.soft_max)
        if step > stride - 2 * self.hard_min and step < stride:
            step = stride - 2 * self.hard_min
        self.soft_max = max(self.soft_max, min(step * 1.1, self.hard_max))
        return step


class LogStats:
    """Running statistics on a logarithmic scale, Welford's algorithm (src/log-stats.h)."""

    def __init__(self):
        self.n, self.m, self.s = 0, 0.0, 0.0

    def add(self, x):
        x = math.log(x)
        self.n += 1
        delta1 = x - self.m
        self.m += delta1 / self.n
        delta2 = x - self.m
        self.s += delta1 * delta2

    def count(self):
        return self.n

    def mean(self):
        return math.exp(self.m)

    def stddev(self):
        return math.sqrt(self.s / self.n)


class Stats:
    """src/time_integration.cpp:41-47."""

    def __init__(self):
        self.good_dt, self.good_dumax, self.bad_dt = LogStats(), LogStats(), LogStats()
        self.max_angle = 0.0

    def report(self):
        """print_stats, src/time_integration.cpp:49-69."""
        out = ["", "Time step statistics:", "",
               "    time steps       count       dt [*]          dumax [*]",
               "    " + "─" * 58]
        ln = "    successful   %9g" % self.good_dt.count()
        if self.good_dt.count():
            ln += "   %8.2e ± %4.2f   %8.2e ± %4.2f" % (
                self.good_dt.mean(), self.good_dt.stddev(), self.good_dumax.mean(), self.good_dumax.stddev())
        out.append(ln)
        ln = "    failed       %9g" % self.bad_dt.count()
        if self.bad_dt.count():
            ln += "   %8.2e ± %4.2f" % (self.bad_dt.mean(), self.bad_dt.stddev())
        out += [ln, "", "    [*] ranges given as (geometric mean) ± (relative stddev)", "",
                "Maximum magnetization angle: %.3g°" % (self.max_angle * 180 / math.pi)]
        return "\n".join(out)


DEFAULT_EVOL_COLUMNS = ("t", "<Mx>", "<My>", "<Mz>", "E_ex", "E_aniso", "E_demag", "E_zeeman", "E_tot")


class Fem:
    """The part of the reference's Fem that surrounds the hot path.

    settings : feellgood_b200.Settings with the loop fields filled in:
        time_step (outputs.evol_time_step), DUMAX, evol_columns, verbose, and the applied field as
        `field`: a callable t -> 3-vector [A/m] (field type RtoR3, Settings::getField) or
        `field_time`: a callable t -> amplitude of mesh.extSpaceField (R4toR3, getFieldTime).
    linAlg   : LinAlgebra (or any object with the same surface; the parity tests drive this very
        loop with the CPU oracle behind the same methods).
    demag    : optional callable(linAlg) run where Fem::compute_all calls myFMM.calc_demag; it
        must write phi/phiv of the NEXT state (LinAlgebra.set_potentials).
    """

    def __init__(self, settings, linAlg, demag=None, region_names=None):
        self.settings, self.linAlg, self.demag = settings, linAlg, demag
        self.region_names = list(region_names or [])
        self.vmax = 0.0
        self.E = np.zeros(4)
        self.Etot0 = math.inf      # avoid "WARNING: energy increased" on first time step
        self.Etot = 0.0
        self.evol = []             # rows of the .evol file
        self.log = []

    # -- settings helpers (Settings::getFieldType / getField / getFieldTime) --
    def _uniform(self):
        return getattr(self.settings, "field_time", None) is None

    def _field(self, t):
        f = getattr(self.settings, "field", None)
        if f is None:
            return np.zeros(3)
        return np.asarray(f(t), dtype=np.float64) if callable(f) else np.asarray(f, dtype=np.float64)

    # -- src/energy.cpp:5-68 --
    def energy(self, t):
        if self._uniform():
            E = self.linAlg.energy(self._field(t))
        else:
            E = self.linAlg.energy(float(self.settings.field_time(t)))
        self.E = np.asarray(E, dtype=np.float64)
        # std::reduce over 4 doubles (libstdc++ random-access path): 0 + ((E0 + E1) + (E2 + E3))
        self.Etot = 0.0 + ((self.E[0] + self.E[1]) + (self.E[2] + self.E[3]))
        if self.settings.verbose and self.Etot > self.Etot0:
            self._say("WARNING: energy increased from %r to %r" % (self.Etot0, self.Etot))

    # -- src/fem.h:159-163 --
    def evolution(self):
        self.linAlg.evolution()
        self.Etot0 = self.Etot

    # -- src/fem.h:203-224 (spin accumulation is outside the path) --
    def compute_all(self, t):
        if self.demag is not None:
            self.demag(self.linAlg)
        self.energy(t)
        self.evolution()

    def _say(self, msg):
        self.log.append(msg)
        if self.settings.verbose:
            print(msg, file=sys.stderr)

    # -- Fem::saver, src/save.cpp:13-140 (the .evol row; .sol files are written by io.py) --
    def saver(self, t_prm, nt):
        cols = getattr(self.settings, "evol_columns", None) or DEFAULT_EVOL_COLUMNS
        la = self.linAlg
        row, cache = [], {}

        def avg(what, region):
            key = (what, region)
            if key not in cache:
                cache[key] = la.avg(what, region)
            return cache[key]
        applied = None
        for col in cols:
            region, key = -1, col
            if ":" in col:
                name, key = col.rsplit(":", 1)
                if name not in self.region_names:
                    raise ValueError("Error: no region named '%s'" % name)
                region = self.region_names.index(name)
            if key == "iter":
                row.append(nt)
            elif key == "t":
                row.append(t_prm.get_t())
            elif key == "dt":
                row.append(t_prm.get_dt())
            elif key == "max_dm":
                row.append(self.vmax * t_prm.get_dt())
            elif key == "max_angle":
                row.append(la.max_angle())
            elif key in ("<Mx>", "<My>", "<Mz>"):
                row.append(avg("u", region)["xyz".index(key[2])])
            elif key in ("<dMx/dt>", "<dMy/dt>", "<dMz/dt>"):
                row.append(avg("v", region)["xyz".index(key[3])])
            elif key in ("E_ex", "E_aniso", "E_demag", "E_zeeman"):
                row.append(self.E[("E_ex", "E_aniso", "E_demag", "E_zeeman").index(key)])
            elif key == "E_tot":
                row.append(self.Etot)
            elif key in ("Hx", "Hy", "Hz"):
                if applied is None:
                    applied = self._field(t_prm.get_t())
                row.append(applied["xyz".index(key[1])])
            else:
                raise ValueError("Error: invalid column name '%s'" % key)
        self.evol.append(row)
        return row

    # -- Fem::time_integration, src/time_integration.cpp:133-245 --
    def time_integration(self, t_prm, should_stop=None):
        """Returns (status, nt).  `should_stop()` stands for exit_if_signal_received."""
        s, la = self.settings, self.linAlg
        self.compute_all(t_prm.get_t())
        flag = 0
        nt_output = 0
        status = 0
        t_initial = t_prm.get_t()
        t_step = s.time_step
        # std::round (src/time_integration.cpp:170): halves away from zero, not Python's banker's rounding
        step_count = round_half_away((t_prm.tf - t_initial) / t_step)
        stepper = TimeStepper(t_prm.get_dt(), t_prm.DTMIN, t_prm.DTMAX)
        stats = self.stats = Stats()
        stats.max_angle = la.max_angle()
        nt = 0
        for step_nb in range(step_count + 1):
            t_target = t_initial + step_nb * t_step
            while t_prm.get_t() < t_target:
                if should_stop is not None and should_stop():
                    self.nt = nt
                    return 1, nt
                t_prm.set_dt(stepper(t_target - t_prm.get_t()))
                last_step = (t_prm.get_dt() == t_target - t_prm.get_t())
                if s.verbose:
                    self._say("-" * 64)
                    if flag:
                        self._say("  TRYING AGAIN with a smaller time step: retry %d" % flag)
                    self._say("evol step = %d, step = %d, t = %r, dt = %r"
                              % (nt_output, nt, t_prm.get_t(), t_prm.get_dt()))
                if t_prm.is_dt_TooSmall():
                    self._say("\n**ABORTED**: dt < DTMIN")
                    status = 1
                    self.nt = nt
                    return status, nt
                la.base_projection()
                if self._uniform():
                    la.prepareElements(self._field(t_prm.get_t()), t_prm)
                else:
                    la.prepareElements(float(s.field_time(t_prm.get_t())), t_prm)
                err = la.solve(t_prm)
                self.vmax = la.get_v_max()
                if err:
                    flag += 1
                    stepper.set_soft_limit(t_prm.get_dt() / 2)
                    stats.bad_dt.add(t_prm.get_dt())
                    continue
                dumax = t_prm.get_dt() * self.vmax
                stats.good_dt.add(t_prm.get_dt())
                stats.good_dumax.add(dumax)
                if s.verbose:
                    self._say("  -> dumax = %r,  vmax = %r" % (dumax, self.vmax))
                stepper.set_soft_limit((s.DUMAX / self.vmax) * 0.95)
                if dumax > s.DUMAX:
                    flag += 1
                    continue
                self.compute_all(t_prm.get_t())
                nt += 1
                flag = 0
                # Prevent rounding errors from making us miss the target.
                if last_step:
                    t_prm.set_t(t_target)
                else:
                    t_prm.inc_t()
                stats.max_angle = max(stats.max_angle, la.max_angle())
            self.saver(t_prm, nt_output)
            nt_output += 1
        self.nt = nt
        return status, nt

    def write_evol(self, path, header=True):
        """The .evol text file: metadata line with the column names (Settings::evolMetadata,
        src/settings.cpp:293-301), then tab-separated rows with 16 significant digits."""
        cols = getattr(self.settings, "evol_columns", None) or DEFAULT_EVOL_COLUMNS
        with open(path, "w") as f:
            if header:
                f.write("## columns: " + "\t".join(cols) + "\n")
            for row in self.evol:
                f.write("\t".join(("%d" % v) if isinstance(v, (int, np.integer)) else ("%.16g" % v)
                                  for v in row) + "\n")
