"""`python -m i2l_cli predict CKPT IMG ...` -- entry point of the `img2latex predict` surface (the package directory
name is not a valid Python identifier, see i2l_import.py)."""
import sys

import i2l_import

if __name__ == "__main__":
    pkg = i2l_import.load()
    from hmer_img2latex_b200.cli import main
    sys.exit(main())
