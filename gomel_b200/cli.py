"""Command-line tools mirroring the reference's cmd/tomel, cmd/towav, cmd/tophase, cmd/fromphase
(cmd/*/main.go): same positional arguments and the same hard-coded configurations, on the GPU path.

    python -m gomel_b200.cli tomel  file.wav          -> file.wav.png      (cmd/tomel/main.go:21-59)
    python -m gomel_b200.cli towav  file.png [rate]   -> file.png.wav      (cmd/towav/main.go:27-44)
    python -m gomel_b200.cli tophase file.wav         -> file.wav.png      (cmd/tophase/main.go:21-55)
    python -m gomel_b200.cli fromphase file.png       -> file.png.wav      (cmd/fromphase/main.go:20-32)
    python -m gomel_b200.cli tomel-dir  in_dir out_dir   every *.wav in one batched GPU call (gomel_b200/batch.py)
    python -m gomel_b200.cli towav-dir  in_dir out_dir   every *.png, grouped by frame count
"""
import sys

from .mel import NewMel
from .phase import Phase


def _mel():
    m = NewMel()                         # cmd/tomel/main.go:24-31, cmd/towav/main.go:30-39
    m.MelFmin, m.MelFmax, m.YReverse = 0, 16000, True
    m.Window, m.NumMels, m.Resolut = 1280, 192, 4096
    m.GriffinLimIterations, m.VolumeBoost = 2, 0.0
    return m


def _phase():
    from .phase import NewPhase
    p = NewPhase()                       # cmd/tophase/main.go:24-27: NumFreqs 768 * 2, YReverse
    p.y_reverse = True
    p.num_freqs = 768 * 2
    return p


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        print(__doc__)
        return 1
    tool, name = argv[0], argv[1]
    try:
        if tool == "tomel":                      # cmd/tomel/main.go:33-59: .flac -> ToMelFlac, .wav -> ToMelWav, else + ".wav"
            if name.endswith(".flac"):
                _mel().ToMelFlac(name, name + ".png")
            else:
                _mel().ToMelWav(name if name.endswith(".wav") else name + ".wav", name + ".png")
        elif tool == "towav":
            m = _mel()
            if len(argv) > 2:
                m.SampleRate = int(argv[2])
            m.ToWavPng(name, name + ".wav")
        elif tool == "tomel-dir":
            from . import batch
            print("\n".join(batch.tomel_dir(name, argv[2], _mel())))
        elif tool == "towav-dir":
            from . import batch
            print("\n".join(batch.towav_dir(name, argv[2], _mel())))
        elif tool == "tophase":                  # cmd/tophase/main.go:29-55 (Go names, Go codec flavour)
            p = _phase()
            if name.endswith(".flac"):
                p.ToPhaseFlac(name, name + ".png")
            else:
                p.ToPhaseWav(name if name.endswith(".wav") else name + ".wav", name + ".png")
        elif tool == "fromphase":                # cmd/fromphase/main.go:20-32
            _phase().ToWavPng(name, name + ".wav")
        else:
            print(__doc__)
            return 1
    except Exception as e:          # the Go mains print the error and exit 1
        print(f"Error: {e}")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
