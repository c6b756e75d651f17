/* Compile-only stand-in: GNU Radio 3.7 hands out boost::shared_ptr; the harness only needs the semantics. */
#pragma once
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
