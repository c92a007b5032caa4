"""Drop-in for the reference's runner.py: same flags, same data/*.p inputs (see phylo_b200/runner.py)."""
from phylo_b200.runner import main

if __name__ == "__main__":
    main()
