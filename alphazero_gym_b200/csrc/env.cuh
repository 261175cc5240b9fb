// env.cuh -- CartPole-v0 / Pendulum-v0 transition functions (gym 0.17.2 / 0.19.0 classic_control,
// third-party to the reference; reached through rl/make_game.py:59-62 and stepped by the search at
// alphazero/search/mcts.py:449 and :686).  Replaces `copy.deepcopy(Env)` + replay from the root
// (mcts.py:443, :680): the hidden state is kept per node and stepped once at expansion.
// All arithmetic is float64 with the rounding sequence of the Python source (built with -fmad=false).
#pragma once
#include "detmath.cuh"

namespace env {

// returns done
__device__ __forceinline__ bool cartpole_step(const double s[4], int action, double o[4], double& reward) {
    const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5, force_mag = 10.0, tau = 0.02;
    const double total_mass = masspole + masscart, polemass_length = masspole * length;
    const double theta_thr = 12 * 2 * 3.141592653589793 / 360, x_thr = 2.4;
    double x = s[0], x_dot = s[1], theta = s[2], theta_dot = s[3];
    const double force = action == 1 ? force_mag : -force_mag;
    double sintheta, costheta;
    det::sincos_(theta, sintheta, costheta);
    const double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
    const double thetaacc =
        (gravity * sintheta - costheta * temp) / (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
    const double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
    x = x + tau * x_dot;
    x_dot = x_dot + tau * xacc;
    theta = theta + tau * theta_dot;
    theta_dot = theta_dot + tau * thetaacc;
    o[0] = x; o[1] = x_dot; o[2] = theta; o[3] = theta_dot;
    reward = 1.0;
    return x < -x_thr || x > x_thr || theta < -theta_thr || theta > theta_thr;
}

// Python float modulo: result takes the sign of the divisor
__device__ __forceinline__ double py_mod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((b < 0) != (r < 0)) r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

// returns done (always false); reward is the raw env reward (-costs)
__device__ __forceinline__ bool pendulum_step(double th, double thdot, float action, double& newth, double& newthdot,
                                              double& reward) {
    const double max_speed = 8, dt = 0.05, g = 10.0, m = 1.0, l = 1.0, pi = 3.141592653589793;
    const float uf = action < -2.0f ? -2.0f : (action > 2.0f ? 2.0f : action);  // np.clip on the f32 action
    const double u = (double)uf;  // numpy-1.x promotion of the f32 torque
    const double an = py_mod(th + pi, 2 * pi) - pi;
    const double costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u);
    double nd = thdot + (-3 * g / (2 * l) * det::sin_(th + pi) + 3.0 / (m * (l * l)) * u) * dt;
    newth = th + nd * dt;
    nd = nd < -max_speed ? -max_speed : (nd > max_speed ? max_speed : nd);
    newthdot = nd;
    reward = -costs;
    return false;
}

__device__ __forceinline__ float4 pendulum_obs(double th, double thdot) {
    double s, c;
    det::sincos_(th, s, c);
    return make_float4((float)c, (float)s, (float)thdot, 0.0f);
}

}  // namespace env
