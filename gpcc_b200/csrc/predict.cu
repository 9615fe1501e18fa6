// placeholder (replaced below in this round): postb / predictions
#include "../../include/gpcc_b200.h"
extern "C" {
int gpcc_postb(gpcc_problem*, const double*, const double*, double, double*, double*) { return -99; }
int gpcc_predict(gpcc_problem*, const double*, const double*, double, const int*, const double*, double*, double*, double*) { return -99; }
int gpcc_predict_loglik(gpcc_problem*, const double*, const double*, double, const int*, const double*, const double*, const double*, double*, int*) { return -99; }
}
