"""TEST INFRASTRUCTURE ONLY -- single-rank stand-in for mpi4py so that the reference's pure-Python
tetragono package (imported from /root/reference, never copied) runs in this container, which has
no MPI library.  Only what tetragono/utility.py touches is provided."""
