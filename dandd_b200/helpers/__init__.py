"""Drop-in counterparts of the reference's helpers/ that sit on the hot path:
only `allpairs` -- the phylogeny glue, plotting and cluster scripts are out of scope."""
