/* -*- c++ -*- ----------------------------------------------------------
   fix B200 -- created by "package b200 ..." (or on demand by -sf b200);
   owns the device context of the B200 package.
------------------------------------------------------------------------- */

#ifdef FIX_CLASS
// clang-format off
FixStyle(B200,FixB200);
// clang-format on
#else

#ifndef LMP_FIX_B200_H
#define LMP_FIX_B200_H

#include "b200_lmp.h"
#include "fix.h"

namespace LAMMPS_NS {

class FixB200 : public Fix {
 public:
  FixB200(class LAMMPS *, int, char **);
  ~FixB200() override;
  int setmask() override;
  void init() override;
  double memory_usage() override;

  b200_ctx *context() { return ctx; }
  int precision() const { return prec; }
  int profile() const { return profile_flag; }
  // turn a negative b200_* status into error->one() with the library's message
  void check(int rc, const char *file, int line);

  // the package fix of this LAMMPS instance; issues "package b200" defaults if absent
  static FixB200 *instance(class LAMMPS *);

 private:
  b200_ctx *ctx;
  int device, prec, profile_flag;
};

}    // namespace LAMMPS_NS

#define B200_CHECK(fixptr, call) (fixptr)->check((call), FLERR)

#endif
#endif
