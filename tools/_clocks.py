"""NVML clock / power / throttle sampler for seconds-long development runs (a thread polling every 100 ms; bench.py
samples from the main thread instead because its steps are short)."""
import statistics
import threading


class Clocks(threading.Thread):
    NAMES = [("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"), ("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
             ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
             ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown")]

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h, self.period = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index), period
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self.rows, self.stop_flag = [], threading.Event()

    def run(self):
        while not self.stop_flag.wait(self.period):
            try:
                self.rows.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                                  self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                  int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:  # noqa: BLE001
                pass

    def mark(self):
        """Index of the next sample: summary(since=mark) summarises one phase of a longer run."""
        return len(self.rows)

    def summary(self, since=0, stop=True):
        if stop:
            self.stop_flag.set(); self.join()
        rows = self.rows[since:]
        if not rows:
            return None
        busy = [r for r in rows if r[1] > 0.5 * max(x[1] for x in rows)] or rows
        return {"sm_mhz_median_under_load": statistics.median(r[0] for r in busy), "sm_max_mhz": self.max_mhz,
                "power_w_max": max(r[1] for r in busy),
                "reasons": [n for n, attr in self.NAMES if any(r[2] & getattr(self.nv, attr, 0) for r in busy)],
                "samples": len(busy)}
