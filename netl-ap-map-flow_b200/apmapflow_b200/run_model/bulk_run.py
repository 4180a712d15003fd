"""``BulkRun``: parameter sweeps over many LCL simulations (apmapflow/run_model/bulk_run.py).

Same public surface as the reference (constructor :57-72, ``dry_run`` :74, ``start`` :107,
``generate_input_files`` :130, ``_combine_run_params`` :208, ``_initialize_run`` :236,
``_check_processes`` :262, ``_start_simulations`` :286).  The scheduling resource changes from CPU
cores and host RAM to GPUs and device memory: ``num_CPUs`` bounds the number of simulations in flight,
``sys_RAM`` is the memory budget in GB of EACH device, and the new ``devices`` keyword lists the GPUs: a
case goes to the device with the most free budget that can hold it (independent cases, no collective).
``spawn_delay`` / ``retest_delay`` are still accepted and honoured when given, but their defaults drop from the
reference's 5 s (time for a child process to claim its RAM, bulk_run.py:262,286) to 0 / 0.05 s: a run is a thread
of this process whose memory is budgeted up front, and a solve takes milliseconds."""
import string
from collections import defaultdict
from itertools import product
from time import sleep

from .. import _get_logger, set_main_logger_level
from .run_model import estimate_req_RAM, run_model

logger = _get_logger(__name__)


class BulkRun(dict):
    def __init__(self, init_input_file, num_CPUs=2, sys_RAM=4.0, **kwargs):
        super().__init__()
        self.init_input_file = init_input_file.clone()
        self.num_CPUs = num_CPUs
        self.sys_RAM = sys_RAM
        self.avail_RAM = sys_RAM * 0.90
        self.input_file_list = []
        self.devices = list(kwargs.pop("devices", [0]))
        self._next_device = 0
        self.update(kwargs)
        logger.debug("Utilizing a maximum of {:d} concurrent simulations, and {:f} gigabtyes of device memory"
                     .format(int(self.num_CPUs), self.sys_RAM))

    def dry_run(self):
        level = logger.getEffectiveLevel()
        set_main_logger_level("debug")
        self._initialize_run()
        logger.info("{:d} simulations would be performed".format(len(self.input_file_list)))
        logger.info("Writing model input files to disk for inspection")
        for inp_file in self.input_file_list:
            inp_file.write_inp_file()
        logger.info("Dry run has been completed.")
        logger.info("Use the start() method to begin simulations")
        set_main_logger_level(level)

    def start(self):
        """Runs every generated input file; like the reference it returns when the QUEUE is empty
        (bulk_run.py:126), the last simulations may still be finishing -- they run on non-daemon threads,
        so a script that ends here still gets them completed (``wait()`` is an extra, not a requirement)."""
        logger.info("Beginning bulk run of simulations")
        self._initialize_run()
        processes, ram_in_use = [], []
        self._start_simulations(processes, ram_in_use, **self)
        while self.input_file_list:
            self._check_processes(processes, ram_in_use, **self)
            self._start_simulations(processes, ram_in_use, **self)
        self._running = processes

    def generate_input_files(self, default_params, default_name_formats, case_identifer="", case_params=None,
                             append=False):
        if case_identifer:
            case_params = case_params or {}
            keys = [k[1] for k in string.Formatter().parse(case_identifer) if k[1]]
            run_cases = []
            for comb in self._combine_run_params({k: default_params[k] for k in keys}):
                ident = case_identifer.format(**comb)
                params = {k: [comb[k]] if k in comb else default_params[k] for k in default_params}
                params.update(case_params.get(ident, {}))
                run_cases.append(params)
        else:
            run_cases = [default_params]
        if not append:
            self.input_file_list = []
        for params in run_cases:
            for comb in self._combine_run_params(params):
                inp_file = self.init_input_file.clone(default_name_formats)
                inp_file.update(comb)
                self.input_file_list.append(inp_file)

    @staticmethod
    def _combine_run_params(run_params):
        run_params = {k: v for k, v in run_params.items() if v}
        return [dict(zip(run_params.keys(), comb)) for comb in product(*run_params.values())]

    def _initialize_run(self):
        logger.info("Assesing memory requirements of each aperture map")
        maps = defaultdict(list)
        for inp_file in self.input_file_list:
            maps[inp_file["APER-MAP"].value].append(inp_file)
        keys = list(maps.keys())
        kw = {k: v for k, v in self.items() if k == "delim"}
        for key, value in zip(keys, estimate_req_RAM(keys, self.avail_RAM, **kw)):
            logger.info("Estimated {:f}gb of device memory reqired for map: {}".format(value, key))
            for inp_file in maps[key]:
                inp_file.RAM_req = value

    @staticmethod
    def _check_processes(processes, RAM_in_use, retest_delay=0.05, **kwargs):
        while True:
            for i, proc in enumerate(processes):
                if proc.poll() is not None:
                    del processes[i]
                    del RAM_in_use[i]
                    return
            sleep(retest_delay)

    def _free_per_device(self, processes, RAM_in_use):
        """Free budget of every device: ``avail_RAM`` is per GPU, a run occupies the device it was sent to."""
        free = {d: self.avail_RAM for d in self.devices}
        for proc, ram in zip(processes, RAM_in_use):
            dev = getattr(proc, "device", self.devices[0])
            free[dev] = free.get(dev, self.avail_RAM) - ram
        return free

    def _start_simulations(self, processes, RAM_in_use, spawn_delay=0.0, **kwargs):
        while len(processes) < self.num_CPUs:
            free = self._free_per_device(processes, RAM_in_use)
            # emptiest device first; ties go round-robin so equal cases still spread over the GPUs
            order = sorted(range(len(self.devices)),
                           key=lambda k: (-free[self.devices[k]], (k - self._next_device) % len(self.devices)))
            device = self.devices[order[0]]
            for i, inp_file in enumerate(self.input_file_list):
                if inp_file.RAM_req <= free[device]:
                    inp_file = self.input_file_list.pop(i)
                    self._next_device = order[0] + 1
                    processes.append(run_model(inp_file, device=device))
                    RAM_in_use.append(inp_file.RAM_req)
                    if spawn_delay > 0:
                        sleep(spawn_delay)
                    break
            else:
                break

    def wait(self):
        """(new) block until the simulations still in flight after ``start()`` have finished."""
        for proc in getattr(self, "_running", []):
            proc.wait()
