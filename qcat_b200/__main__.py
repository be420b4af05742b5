"""`python -m qcat_b200 [--devices 0,1 | --devices all] <qcat arguments>`: the reference's own command line
(qcat/cli.py, unchanged) on top of the drop-in.

Needs an installed nanoporetech/qcat: its argument parser, logging, batching and writers stay in charge; only the
scanners' detection entry points are rebound to the CUDA path (qcat_b200.dropin).  Everything after the optional
--devices option is handed to qcat.cli.main as it is."""
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    devices = None
    if argv[:1] == ["--devices"] and len(argv) >= 2:
        devices = "all" if argv[1] == "all" else [int(v) for v in argv[1].split(",")]
        argv = argv[2:]
    try:
        import qcat.scanner  # noqa: F401  (import order: scanner <-> scanner_epi2me cycle)
        from qcat import cli
    except ImportError as exc:
        sys.stderr.write("qcat_b200: nanoporetech/qcat is not importable (%s); install it, or use qcat_b200.scanner / "
                         "qcat_b200.fastx.demux_file directly\n" % exc)
        return 2
    from qcat_b200 import dropin
    dropin.install(devices=devices)
    try:
        return cli.main(argv)
    finally:
        dropin.uninstall()


if __name__ == "__main__":
    sys.exit(main() or 0)
