"""CPU tests of the drop-in command line: flag surface, validation order and exit codes of
cudabrot.cu:579-754 (nothing here reaches the GPU; argument errors exit before SetupCUDA)."""
import subprocess

import pytest


@pytest.fixture(scope="module")
def cli(buddha):
    return buddha.capi.CLI_PATH


def run(cli, *args):
    return subprocess.run([cli] + list(args), capture_output=True, text=True, timeout=60)


def test_help_prints_usage_and_exits_zero(cli):
    r = run(cli, "--help")
    assert r.returncode == 0                      # PrintUsage -> exit(0), cudabrot.cu:619
    assert r.stdout.startswith("Usage: ")
    for flag in ["--help", "-d <device number>", "-o <output file name>", "-m <max escape",
                 "-c <min escape", "-g <gamma", "-t <seconds", "-w <width>", "-h <height>",
                 "-s <save/load file>", "--min-real", "--max-real", "--min-imag", "--max-imag"]:
        assert flag in r.stdout


@pytest.mark.parametrize("args,message", [
    (["-w", "0"], "Output width must be positive."),
    (["-h", "-4"], "Output height must be positive."),
    (["--max-real", "-2"], "Maximum real value must be greater than minimum real value."),
    # flag-order dependence: 2.0 <= 3.0 at the first flag
    (["--min-real", "3", "--max-real", "4"],
     "Maximum real value must be greater than minimum real value."),
    (["--max-imag", "-3"], "Minimum imaginary value must be greater than maximum imaginary value."),
    (["-w"], "Argument -w needs a value."),
    (["-w", "12x"], "Invalid number given to argument -w: 12x"),
    (["-g", ""], "Invalid number given to argument -g: "),
    (["-o"], "Missing output file name."),
    (["-s"], "Missing in-progress buffer file name."),
    (["--frobnicate"], "Invalid argument: --frobnicate"),
])
def test_invalid_arguments_print_message_usage_and_exit_zero(cli, args, message):
    r = run(cli, *args)
    assert r.returncode == 0
    lines = r.stdout.splitlines()
    assert lines[0] == message
    assert lines[1].startswith("Usage: ")


def test_valid_order_is_accepted_and_high_m_warns(cli):
    # --max-real first makes the same canvas valid; then the run dies at device creation here
    r = run(cli, "--max-real", "4", "--min-real", "3", "-m", "60001", "-t", "0")
    assert "Warning: Using a high number of iterations" in r.stdout
    assert "Creating 1000x1000 image, 60001 max iterations." in r.stdout
