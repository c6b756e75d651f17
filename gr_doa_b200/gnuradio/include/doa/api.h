/* Symbol visibility of the gnuradio-doa library (same macro name as gr-doa's include/doa/api.h:27-31). */
#ifndef INCLUDED_DOA_API_H
#define INCLUDED_DOA_API_H
#include <gnuradio/attributes.h>
#ifdef gnuradio_doa_EXPORTS
#define DOA_API __GR_ATTR_EXPORT
#else
#define DOA_API __GR_ATTR_IMPORT
#endif
#endif
