/* -*- c++ -*- */
/* SWIG interface of the four GPU-backed blocks (GNU Radio 3.7: GR_SWIG_BLOCK_MAGIC2), same python names as gr-doa
   swig/doa_swig.i:22-36.  The remaining gr-doa blocks (antenna_correction, calibrate_lin_array, Connex variants) are not
   part of this hot path; keep their lines from the original file when merging. */
/* Notes for whoever merges this into gr-doa's swig/doa_swig.i:
   - every block listed here is backed by libdoa_cuda (link gnuradio-doa against it, lib/CMakeLists.snippet.txt); the python
     names, constructor arguments and port shapes are the reference's, so apps/*.py and apps/*.grc keep working;
   - autocorrelate gains one method, set_antenna_config(filename), which SWIG exposes automatically from the header;
   - music_chain is new (autocorrelate -> MUSIC_lin_array -> find_local_max in one GPU call); its second factory,
     music_chain::make_sc16 (UHD "sc16" input items), reaches python as the flat SWIG name doa.music_chain_make_sc16
     (GR_SWIG_BLOCK_MAGIC2 rebinds doa.music_chain to make(); grc/doa_music_chain_sc16.xml uses the flat name);
   - GR_SWIG_BLOCK_MAGIC2 needs the headers both in the %{ %} block (for the wrapper's C++) and as %include (for SWIG). */
#define DOA_API
%include "gnuradio.i"
%{
#include "doa/autocorrelate.h"
#include "doa/MUSIC_lin_array.h"
#include "doa/rootMUSIC_linear_array.h"
#include "doa/find_local_max.h"
#include "doa/music_chain.h"
#include "doa/rootmusic_chain.h"
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
%include "doa/rootmusic_chain.h"
GR_SWIG_BLOCK_MAGIC2(doa, rootmusic_chain);   /* not in gr-doa: autocorrelate + rootMUSIC_linear_array in one GPU call */
%include "doa/calibrate_lin_array.h"
GR_SWIG_BLOCK_MAGIC2(doa, calibrate_lin_array);
