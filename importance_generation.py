"""Drop-in entry point: same name and flags as the reference's importance_generation.py."""
from dct_pruning_b200.cli import main

if __name__ == '__main__':
    main()
