/* -*- c++ -*- */
/* SWIG interface of the four GPU-backed blocks (GNU Radio 3.7: GR_SWIG_BLOCK_MAGIC2), same python names as gr-doa
   swig/doa_swig.i:22-36.  The remaining gr-doa blocks (antenna_correction, calibrate_lin_array, Connex variants) are not
   part of this hot path; keep their lines from the original file when merging. */
#define DOA_API
%include "gnuradio.i"
%{
#include "doa/autocorrelate.h"
#include "doa/MUSIC_lin_array.h"
#include "doa/rootMUSIC_linear_array.h"
#include "doa/find_local_max.h"
#include "doa/music_chain.h"
#include "doa/calibrate_lin_array.h"
%}
%include "doa/autocorrelate.h"
GR_SWIG_BLOCK_MAGIC2(doa, autocorrelate);
%include "doa/MUSIC_lin_array.h"
GR_SWIG_BLOCK_MAGIC2(doa, MUSIC_lin_array);
%include "doa/rootMUSIC_linear_array.h"
GR_SWIG_BLOCK_MAGIC2(doa, rootMUSIC_linear_array);
%include "doa/find_local_max.h"
GR_SWIG_BLOCK_MAGIC2(doa, find_local_max);
%include "doa/music_chain.h"
GR_SWIG_BLOCK_MAGIC2(doa, music_chain);   /* not in gr-doa: the three blocks above in one GPU call */
%include "doa/calibrate_lin_array.h"
GR_SWIG_BLOCK_MAGIC2(doa, calibrate_lin_array);
